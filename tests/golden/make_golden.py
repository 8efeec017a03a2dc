"""Generate golden vectors by running the REFERENCE itself (bioframe v0.8.0
from /root/reference/src, imported with a stub `matplotlib`).

Run in the build container only:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as
small .npz fixtures next to this script.  Inputs are stored with the outputs.

Closest cases use targets with unique starts and unique ends: the reference
orders tied targets with an unstable argsort (arrops.py:562,580), so only
tie-free keys have a defined answer.
"""
from __future__ import annotations

import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    sys.dont_write_bytecode = True
    stub = tempfile.mkdtemp(prefix="mplstub")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for name, body in (("__init__", ""), ("pyplot", ""), ("colors", "def to_rgb(c):\n    return (0.0, 0.0, 0.0)\n")):
        with open(os.path.join(stub, "matplotlib", name + ".py"), "w") as f:
            f.write(body)
    sys.path[:0] = [stub, "/root/reference/src"]
    import bioframe  # noqa: E402

    return bioframe


def ustr(col):
    return np.array([str(x) for x in col], dtype="U16")


def tie_free(rng, n, lo, hi, maxlen):
    """n intervals with unique starts and unique ends."""
    starts = rng.choice(np.arange(lo, hi), size=n, replace=False).astype(np.int64)
    ends = np.empty(n, dtype=np.int64)
    used = set()
    for t in range(n):
        e = int(starts[t] + rng.integers(0, maxlen))
        while e in used:
            e += 1
        used.add(e)
        ends[t] = e
    return starts, ends


def rand_set(rng, n, lo, hi, maxlen, p_point=0.15):
    s = rng.integers(lo, hi, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return s, s + ln


def arrops_cases(bf, out):
    ar = bf.core.arrops
    rng = np.random.default_rng(20261019)
    cases = {
        "doc": (np.array([1, 3, 8, 12]), np.array([5, 8, 10, 14]), np.array([4, 10]), np.array([8, 11])),
        "mixed": (
            np.array([5, 5, 5, 1, 7, 7, 30]),
            np.array([9, 7, 9, 20, 7, 8, 31]),
            np.array([5, 6, 5, 7, 0, 19, 31]),
            np.array([6, 8, 6, 7, 40, 20, 31]),
        ),
        "empty2": (np.array([1, 5]), np.array([3, 9]), np.array([], dtype=np.int64), np.array([], dtype=np.int64)),
        "empty1": (np.array([], dtype=np.int64), np.array([], dtype=np.int64), np.array([1, 5]), np.array([3, 9])),
    }
    for t in range(10):
        n1, n2 = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        lo, hi = (-500, 500) if t % 2 else (0, 3000)
        s1, e1 = rand_set(rng, n1, lo, hi, 40 + 30 * t)
        s2, e2 = rand_set(rng, n2, lo, hi, 40 + 30 * t)
        cases[f"rand{t}"] = (s1, e1, s2, e2)
    # a wide-coordinate case: negative starts and INT64_MAX ends (complement/subtract inputs)
    big = np.iinfo(np.int64).max
    cases["wide"] = (
        np.array([-7, 0, 10, 100, 5]),
        np.array([big, 50, big, 200, 5]),
        np.array([-100, 20, 150, big - 5, 0]),
        np.array([-50, 30, big, big, 0]),
    )
    names = []
    for name, (s1, e1, s2, e2) in cases.items():
        s1, e1, s2, e2 = (np.asarray(a, dtype=np.int64) for a in (s1, e1, s2, e2))
        names.append(name)
        for key, val in (("s1", s1), ("e1", e1), ("s2", s2), ("e2", e2)):
            out[f"{name}/{key}"] = val
        for closed in (False, True):
            i, j = ar.overlap_intervals(s1, e1, s2, e2, closed=closed)
            out[f"{name}/ov{int(closed)}_1"], out[f"{name}/ov{int(closed)}_2"] = i, j
        i, j, u1, u2 = ar.overlap_intervals_outer(s1, e1, s2, e2)
        out[f"{name}/outer_u1"], out[f"{name}/outer_u2"] = u1, u2
        for tag, md in (("0", 0), ("none", None), ("2", 2), ("75", 75)):
            if len(s1) == 0:
                continue
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ids, cs, ce = ar.merge_intervals(s1, e1, min_dist=md)
            out[f"{name}/merge_{tag}_ids"] = ids
            out[f"{name}/merge_{tag}_s"] = cs
            out[f"{name}/merge_{tag}_e"] = ce
        if len(s1):
            for tag, b in (("def", (0, big)), ("b", (int(s1.min()) + 3, int(e1.max()) - 3)), ("out", (int(s1.min()) - 9, int(min(e1.max(), big - 9)) + 9))):
                cs, ce = ar.complement_intervals(s1, e1, bounds=b)
                out[f"{name}/compl_{tag}_bounds"] = np.array(b, dtype=np.int64)
                out[f"{name}/compl_{tag}_s"] = cs
                out[f"{name}/compl_{tag}_e"] = ce
    out["names"] = np.array(names)

    # closest: tie-free targets
    cl_names = []
    for t in range(8):
        n1, n2 = int(rng.integers(1, 300)), int(rng.integers(2, 300))
        s1, e1 = rand_set(rng, n1, 0, 4000, 60)
        s2, e2 = tie_free(rng, n2, 0, 4000, 60)
        along = rng.random(n1) < 0.6
        name = f"cl{t}"
        cl_names.append(name)
        for key, val in (("s1", s1), ("e1", e1), ("s2", s2), ("e2", e2), ("along", along)):
            out[f"{name}/{key}"] = val
        variants = {
            "k1": dict(k=1),
            "k2": dict(k=2),
            "k5": dict(k=5),
            "k1_noov": dict(k=1, ignore_overlaps=True),
            "k3_noov": dict(k=3, ignore_overlaps=True),
            "k2_noup": dict(k=2, ignore_upstream=True),
            "k2_nodn": dict(k=2, ignore_downstream=True),
            "k2_noov_noup": dict(k=2, ignore_overlaps=True, ignore_upstream=True),
            "k2_along": dict(k=2, along=along),
            "k2_along_noup": dict(k=2, along=along, ignore_upstream=True),
            "k3_along_nodn_noov": dict(k=3, along=along, ignore_downstream=True, ignore_overlaps=True),
        }
        for tag, kw in variants.items():
            i, j = ar.closest_intervals(s1, e1, s2, e2, **kw)
            out[f"{name}/{tag}_1"], out[f"{name}/{tag}_2"] = i, j
        # self-closest on the tie-free set
        for tag, kw in {"self_k1": dict(k=1), "self_k3": dict(k=3)}.items():
            i, j = ar.closest_intervals(s2, e2, None, None, **kw)
            out[f"{name}/{tag}_1"], out[f"{name}/{tag}_2"] = i, j
    out["closest_names"] = np.array(cl_names)


def frame_cases(bf, out):
    from bioframe_b200 import synth

    names = np.array(synth.HG38_CHROMS + ["chrM", "chrUn"], dtype=object)
    rng = np.random.default_rng(7)

    def frame(n, seed, extra_code, tie_free_targets=False):
        c, s, e = synth.make_intervals(n, seed, "geom1k")
        # shrink coordinates so that overlaps are plentiful at this size
        s = s // 2000
        e = s + rng.integers(0, 40, n)  # lengths 0..39 (points included)
        c = c.copy()
        c[rng.random(n) < 0.02] = extra_code  # a chromosome only this frame has
        if tie_free_targets:
            # unique (chrom, start) and unique (chrom, end)
            key = c.astype(np.int64) * (1 << 40)
            _, first = np.unique(key + s, return_index=True)
            c, s, e = c[first], s[first], e[first]
            key = c.astype(np.int64) * (1 << 40)
            order = np.argsort(key + e, kind="stable")
            c, s, e = c[order], s[order], e[order]
            bump = np.zeros(len(e), dtype=np.int64)
            for t in range(1, len(e)):
                if c[t] == c[t - 1] and e[t] + bump[t] <= e[t - 1] + bump[t - 1]:
                    bump[t] = e[t - 1] + bump[t - 1] + 1 - e[t]
            e = e + bump
            perm = rng.permutation(len(e))
            c, s, e = c[perm], s[perm], e[perm]
        strand = np.where(rng.random(len(c)) < 0.5, "+", "-")
        return pd.DataFrame({"chrom": names[c], "start": s, "end": e, "strand": strand})

    df1 = frame(20000, 1, 24)
    df2 = frame(15000, 2, 25)
    df2u = frame(15000, 3, 25, tie_free_targets=True)
    for tag, df in (("df1", df1), ("df2", df2), ("df2u", df2u)):
        out[f"{tag}/chrom"] = ustr(df["chrom"])
        out[f"{tag}/start"] = df["start"].values
        out[f"{tag}/end"] = df["end"].values
        out[f"{tag}/strand"] = ustr(df["strand"])

    # record the set order this process used, so a test can replay it
    g1 = df1.groupby(["chrom"], observed=True, dropna=False).indices
    g2 = df2.groupby(["chrom"], observed=True, dropna=False).indices
    order = [k[0] if isinstance(k, tuple) else k for k in set.union(set(g1), set(g2))]
    out["overlap_group_order"] = np.array(order)
    for how in ("inner", "left", "right", "outer"):
        i, j = bf.ops._overlap_intidxs(df1, df2, how=how)
        out[f"overlap_{how}_1"], out[f"overlap_{how}_2"] = i, j

    g1 = df1.groupby("chrom", observed=True).indices
    g2 = df2u.groupby("chrom", observed=True).indices
    out["closest_group_order"] = np.array(list(set.union(set(g1), set(g2))))
    variants = {
        "k1": dict(k=1),
        "k3": dict(k=3),
        "k2_noov": dict(k=2, ignore_overlaps=True),
        "k2_strand_noup": dict(k=2, direction_col="strand", ignore_upstream=True),
        "k2_strand_nodn": dict(k=2, direction_col="strand", ignore_downstream=True),
    }
    for tag, kw in variants.items():
        i, j = bf.ops._closest_intidxs(df1, df2u, **kw)
        out[f"closest_{tag}_1"], out[f"closest_{tag}_2"] = i, j

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tag, md in (("0", 0), ("none", None), ("25", 25)):
            cl = bf.cluster(df1, min_dist=md)
            out[f"cluster_{tag}_ids"] = cl["cluster"].values
            out[f"cluster_{tag}_cs"] = cl["cluster_start"].values
            out[f"cluster_{tag}_ce"] = cl["cluster_end"].values
            mg = bf.merge(df1, min_dist=md)
            out[f"merge_{tag}_chrom"] = ustr(mg["chrom"])
            out[f"merge_{tag}_s"] = mg["start"].values
            out[f"merge_{tag}_e"] = mg["end"].values
            out[f"merge_{tag}_n"] = mg["n_intervals"].values
        cv = bf.coverage(df1.copy(), df2)
        out["coverage"] = cv["coverage"].values
        cp = bf.complement(df1[["chrom", "start", "end"]])
        out["complement_chrom"] = ustr(cp["chrom"])
        out["complement_s"] = cp["start"].values
        out["complement_e"] = cp["end"].values
        view = pd.DataFrame(
            {
                "chrom": ["chr1", "chr1", "chr2", "chr9", "chrM"],
                "start": [0, 60000, 100, 0, 0],
                "end": [50000, 124000, 90000, 70000, 16569],
                "name": ["p", "q", "b", "a", "m"],
            }
        )
        cp = bf.complement(df1[["chrom", "start", "end"]], view)
        out["complement_view_chrom"] = ustr(cp["chrom"])
        out["complement_view_s"] = cp["start"].values
        out["complement_view_e"] = cp["end"].values
        out["complement_view_region"] = ustr(cp["view_region"])


def main():
    bf = import_reference()
    a = {}
    arrops_cases(bf, a)
    np.savez_compressed(os.path.join(HERE, "arrops_cases.npz"), **a)
    f = {}
    frame_cases(bf, f)
    np.savez_compressed(os.path.join(HERE, "frame_cases.npz"), **f)
    print("arrops keys", len(a), "frame keys", len(f))
    for fn in ("arrops_cases.npz", "frame_cases.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
