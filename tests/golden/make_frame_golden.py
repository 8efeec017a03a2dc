"""Frame-level golden outputs from the REFERENCE itself (bioframe v0.8.0 imported
from /root/reference/src with a stub matplotlib).  Run in the build container only:

    python tests/golden/make_frame_golden.py

Writes tests/golden/frames.pkl: {case name: dict(args..., expected DataFrame)}.
The reference cannot travel to the GPU box; these fixtures do.  `overlap`
(how != 'left') and `closest` visit chromosomes in `set` order, which depends on
PYTHONHASHSEED: the tests compare those frames after a canonical row sort (the
exact block order is pinned at the index level by frame_cases.npz)."""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def frames(rng, n1=60, n2=45, with_na=False, categorical=False, tie_free=False):
    """tie_free: unique starts and unique ends (closest picks among targets with equal end / equal
    start through an unstable argsort in the reference, so only tie-free inputs have a defined answer)"""
    chroms = np.array(["chr1", "chr2", "chrX", "chr10"], dtype=object)

    def one(n, only=None):
        c = rng.choice(chroms[:3] if only is None else only, n)
        if tie_free:
            s = rng.choice(np.arange(0, 4000, 7), n, replace=False)  # multiples of 7
            ln = 1 + 7 * rng.integers(0, 6, n) + rng.integers(0, 6, n)  # end = 7a + r, r in 1..6 -> unique per start
            e = s + ln
            while len(set(e)) < n:  # re-draw colliding ends
                dup = pd.Series(e).duplicated().values
                e[dup] = s[dup] + 1 + rng.integers(0, 60, dup.sum())
            ln = e - s
        else:
            s = rng.integers(0, 400, n)
            ln = rng.integers(0, 40, n)
        strand = rng.choice(np.array(["+", "-", "."], dtype=object), n)
        return pd.DataFrame({"chrom": c, "start": s, "end": s + ln, "strand": strand, "score": rng.integers(0, 9, n)})

    df1, df2 = one(n1), one(n2, only=chroms[1:])
    if with_na:
        for df in (df1, df2):
            df["start"] = df["start"].astype(pd.Int64Dtype())
            df["end"] = df["end"].astype(pd.Int64Dtype())
            df.loc[df.index[::17], ["chrom", "start", "end"]] = pd.NA
    if categorical:
        df1["chrom"] = pd.Categorical(df1["chrom"], categories=["chrX", "chr2", "chr1", "chr10"])
        df2["chrom"] = pd.Categorical(df2["chrom"], categories=["chr10", "chrX", "chr2", "chr1"])
    return df1, df2


def main():
    bf = import_reference()
    pd.set_option("future.infer_string", False)
    rng = np.random.default_rng(2024)
    cases = {}

    def add(name, fn, **kw):
        cases[name] = dict(fn=fn, kwargs=kw, expected=getattr(bf, fn)(**{k: (v.copy() if isinstance(v, pd.DataFrame) else v) for k, v in kw.items()}))

    for tag, opts in (("plain", {}), ("na", dict(with_na=True)), ("cat", dict(categorical=True))):
        df1, df2 = frames(rng, **opts)
        for how in ("left", "inner", "right", "outer"):
            add(f"overlap_{tag}_{how}", "overlap", df1=df1, df2=df2, how=how, return_index=True, return_overlap=True)
        add(f"overlap_{tag}_left_noorder", "overlap", df1=df1, df2=df2, how="left", keep_order=False, suffixes=("_a", "_b"))
        if tag != "na":
            add(f"overlap_{tag}_on", "overlap", df1=df1, df2=df2, how="outer", on=["strand"], return_index=True)
        add(f"cluster_{tag}", "cluster", df=df1)
        add(f"cluster_{tag}_d5", "cluster", df=df1, min_dist=5)
        add(f"cluster_{tag}_none", "cluster", df=df1, min_dist=None, return_input=False)
        add(f"merge_{tag}", "merge", df=df1)
        add(f"merge_{tag}_d7", "merge", df=df1, min_dist=7)
        if tag != "na":
            add(f"cluster_{tag}_on", "cluster", df=df1, on=["strand"])
            add(f"merge_{tag}_on", "merge", df=df1, on=["strand"], min_dist=None)
        add(f"coverage_{tag}", "coverage", df1=df1, df2=df2)
        add(f"count_overlaps_{tag}", "count_overlaps", df1=df1, df2=df2)
        add(f"setdiff_{tag}", "setdiff", df1=df1, df2=df2)
        if tag != "na":
            add(f"count_overlaps_{tag}_on", "count_overlaps", df1=df1, df2=df2, on=["strand"], return_input=False)
            add(f"setdiff_{tag}_on", "setdiff", df1=df1, df2=df2, on=["strand"])
        df1, df2 = frames(rng, tie_free=True, **opts)
        add(f"closest_{tag}_k1", "closest", df1=df1, df2=df2)
        add(f"closest_{tag}_k3", "closest", df1=df1, df2=df2, k=3, return_index=True, return_overlap=True)
        add(f"closest_{tag}_noov", "closest", df1=df1, df2=df2, k=2, ignore_overlaps=True)
        if tag != "na":
            add(f"closest_{tag}_strand_up", "closest", df1=df1, df2=df2, k=2, direction_col="strand", ignore_upstream=True)
            add(f"closest_{tag}_strand_dn", "closest", df1=df1, df2=df2, k=2, direction_col="strand", ignore_downstream=True)
        add(f"closest_{tag}_self", "closest", df1=df1, df2=None, k=2)
        add(f"complement_{tag}", "complement", df=df1)
        view = pd.DataFrame({"chrom": ["chr1", "chr1", "chr2", "chrX", "chr10"], "start": [0, 150, 10, 0, 0],
                             "end": [120, 500, 300, 350, 80], "name": ["p", "q", "chr2", "x", "ten"]})
        add(f"complement_{tag}_view", "complement", df=df1, view_df=view)
        add(f"complement_{tag}_dict", "complement", df=df1, view_df={"chr1": 500, "chr2": 450, "chrX": 10**6, "chr10": 77})
    # subtract: drawn from a generator of its own so that the cases above keep their inputs
    rng2 = np.random.default_rng(77)
    for tag, opts in (("plain", {}), ("na", dict(with_na=True)), ("cat", dict(categorical=True))):
        df1, df2 = frames(rng2, **opts)
        add(f"subtract_{tag}", "subtract", df1=df1, df2=df2)
        add(f"subtract_{tag}_index", "subtract", df1=df1, df2=df2, return_index=True, suffixes=("_a", "_b"))
        shuffled = df1.copy()
        shuffled.index = rng2.permutation(len(shuffled))  # labels do not order like rows: composed path
        add(f"subtract_{tag}_shuffled", "subtract", df1=shuffled, df2=df2, return_index=True)
        if tag != "na":
            dense = pd.concat([df2, df2.assign(start=df2["start"] // 2, end=df2["end"] // 2 + 30)], ignore_index=True)
            add(f"subtract_{tag}_dense", "subtract", df1=df1, df2=dense, return_index=True)
    edge1 = pd.DataFrame({"chrom": ["chr1"] * 8 + ["chr3"], "start": [0, 0, 5, 5, 10, 12, 30, 3, 7],
                          "end": [0, 8, 5, 10, 20, 12, 31, 40, 9], "name": list("abcdefghi")})
    edge2 = pd.DataFrame({"chrom": ["chr1"] * 6 + ["chr2"], "start": [0, 5, 8, 10, 12, 30, 1], "end": [3, 5, 10, 11, 15, 30, 2]})
    add("subtract_edges", "subtract", df1=edge1, df2=edge2, return_index=True)
    add("subtract_edges_cols", "subtract", df1=edge1.rename(columns={"chrom": "c", "start": "s", "end": "e"})[["name", "c", "s", "e"]],
        df2=edge2, cols1=("c", "s", "e"))
    add("subtract_no_targets", "subtract", df1=edge1, df2=edge2.iloc[:0])
    # assign_view / trim / pair_by_distance: frame glue over the device overlap
    view3 = pd.DataFrame({"chrom": ["chr1", "chr1", "chr2", "chrX"], "start": [0, 160, 20, 0], "end": [150, 333, 410, 90],
                          "name": ["left", "right", "two", "x"]})
    for tag, opts in (("plain", {}), ("cat", dict(categorical=True))):
        df1, _ = frames(rng2, **opts)
        add(f"assign_view_{tag}", "assign_view", df=df1, view_df=view3)
        add(f"assign_view_{tag}_drop", "assign_view", df=df1, view_df=view3, drop_unassigned=True, df_view_col="reg")
        add(f"trim_{tag}_view", "trim", df=df1, view_df=view3)
        add(f"trim_{tag}_noview", "trim", df=df1.assign(start=df1["start"] - 30))
        named = df1.assign(region=np.where(df1["chrom"].astype(str) == "chr2", "two", None))
        add(f"trim_{tag}_col", "trim", df=named, view_df=view3, df_view_col="region", return_view_columns=True)
        one = df1[df1["chrom"].astype(str) == "chr1"].copy()
        add(f"pair_by_distance_{tag}", "pair_by_distance", df=one, min_sep=5, max_sep=90)
        add(f"pair_by_distance_{tag}_mid", "pair_by_distance", df=df1, min_sep=0, max_sep=60, min_intervening=1,
            max_intervening=6, return_index=True, keep_order=True)
        add(f"pair_by_distance_{tag}_ends", "pair_by_distance", df=df1, min_sep=10, max_sep=200, relative_to="endpoints",
            suffixes=("_a", "_b"))
    # mark_runs / merge_runs: bedGraph-like input (non-overlapping bins with a value column)
    rng3 = np.random.default_rng(5)
    bins = []
    for chrom in ("chr2", "chr1", "chrX"):
        pos = 0
        for _ in range(40):
            pos += int(rng3.integers(0, 3)) * 5  # some gaps
            width = int(rng3.integers(1, 20))
            bins.append((chrom, pos, pos + width, int(rng3.integers(0, 3)), float(rng3.integers(0, 2)) / 2))
            pos += width
    graph = pd.DataFrame(bins, columns=["chrom", "start", "end", "state", "signal"])
    graph = graph.sample(frac=1.0, random_state=3)  # shuffled rows, labels kept
    add("mark_runs_int", "mark_runs", df=graph, col="state")
    add("mark_runs_float_continue", "mark_runs", df=graph, col="signal", reset_counter=False, run_col="r")
    add("merge_runs_int", "merge_runs", df=graph, col="state")
    add("merge_runs_agg", "merge_runs", df=graph, col="signal", agg={"n": ("state", "count"), "top": ("state", "max")})
    add("mark_runs_overlapping", "mark_runs", df=frames(rng3)[0], col="score", allow_overlaps=True)
    # nulls inside an `on` column: groupby(dropna=False) of the reference makes ('chr1', nan) ONE key
    # shared by both frames (ops.py:287-288), so such rows are overlapped with each other; all-null rows
    # (null chromosome) still never pair.  Own generator: the cases above keep their inputs.
    rng4 = np.random.default_rng(404)
    for tag, opts in (("na", dict(with_na=True)), ("plain", {})):
        df1, df2 = frames(rng4, n1=70, n2=55, **opts)
        for df in (df1, df2):
            df.loc[df.index[3::5], "strand"] = None
        for how in ("inner", "left", "outer", "right"):
            add(f"overlap_{tag}_on_null_{how}", "overlap", df1=df1, df2=df2, how=how, on=["strand"], return_index=True)
        add(f"count_overlaps_{tag}_on_null", "count_overlaps", df1=df1, df2=df2, on=["strand"])
        add(f"setdiff_{tag}_on_null", "setdiff", df1=df1, df2=df2, on=["strand"])
    with open(os.path.join(HERE, "frames.pkl"), "wb") as f:
        pickle.dump(cases, f, protocol=4)
    print(len(cases), "cases ->", os.path.join(HERE, "frames.pkl"))


if __name__ == "__main__":
    main()
