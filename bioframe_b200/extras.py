"""Operations the reference layers on top of `overlap` and the running-max scan (src/bioframe/extras.py) --
SURVEY.md 8f "next" #3: `pair_by_distance` (extras.py:389-543), `mark_runs` (extras.py:546-652) and
`merge_runs` (extras.py:655-735)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import ops
from .core import checks
from .core.specs import _get_default_colnames

__all__ = ["pair_by_distance", "mark_runs", "merge_runs"]


def pair_by_distance(df, min_sep, max_sep, min_intervening=None, max_intervening=None, relative_to="midpoints",
                     cols=None, return_index=False, keep_order=False, suffixes=("_1", "_2")):
    """All pairs of intervals of `df` whose separation lies in [min_sep, max_sep] bp, optionally limited by the
    number of intervals between them (extras.py:389-543).  Every interval throws a probe to its right and one
    to its left; a pair qualifies when the right probe of the upstream interval meets the left probe of the
    downstream one -- an inner overlap of the two probe sets, enumerated on the device."""
    df = df.copy()
    stem = return_index if isinstance(return_index, str) else "index"
    icol1, icol2 = stem + suffixes[0], stem + suffixes[1]
    if return_index or keep_order:
        df.index.name = stem
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    df = df.sort_values([ck, sk, ek]).reset_index(drop=False)

    if min_sep >= max_sep:
        raise ValueError("min_sep must be less than max_sep")
    if min_sep < 0:
        raise ValueError("min_sep must be >=0")
    if min_intervening is None:
        min_intervening = 0
    if max_intervening is None:
        max_intervening = df.index.max()
    if min_intervening > max_intervening:
        raise ValueError("min_intervening must be less or equal to max_intervening")
    if min_intervening < 0:
        raise ValueError("min_intervening must be >=0")
    if relative_to not in ("midpoints", "endpoints"):
        raise ValueError("relative_to must either specify 'midpoints' or 'endpoints' ")

    mids = (df[sk] + df[ek]) // 2
    anchor_right = df[ek] if relative_to == "endpoints" else mids  # where the probe to the right starts from
    anchor_left = df[sk] if relative_to == "endpoints" else mids
    carried = [ck, stem] if (return_index or keep_order) else [ck]
    right_probe = df[carried].copy()
    right_probe[sk] = anchor_right + min_sep // 2
    right_probe[ek] = anchor_right + (max_sep + 1) // 2
    left_probe = df[carried].copy()
    left_probe[sk] = anchor_left - max_sep // 2
    left_probe[ek] = anchor_left - (min_sep + 1) // 2

    pairs = ops.overlap(right_probe, left_probe, suffixes=suffixes, how="inner", return_index=True, return_input=False)
    between = np.abs(pairs[f"index{suffixes[0]}"] - pairs[f"index{suffixes[1]}"]) - 1
    pairs = pairs[(between <= max_intervening) & (between >= min_intervening)]

    def side(suffix):
        rows = df.iloc[pairs[f"index{suffix}"].values]
        return rows.rename(columns=lambda c: c + suffix).reset_index(drop=True)

    out = pd.concat([side(suffixes[0]), side(suffixes[1])], axis=1)
    if keep_order:
        out = out.sort_values([icol1, icol2])
    if not return_index:
        out = out.drop([icol1, icol2], axis=1)
    out.reset_index(drop=True, inplace=True)
    return out


def mark_runs(df, col, *, allow_overlaps=False, reset_counter=True, run_col="run", cols=None):
    """Number the runs of spatially consecutive intervals that share the value of `col` (extras.py:546-652).
    Returns a reordered copy of `df`: chromosomes in order of first appearance, rows by (start, end), plus
    `run_col`.

    The reference loops over chromosomes: sort, running max of the ends, a border wherever an interval starts
    past everything before it or the value changes, cumulative sum.  Here the spatial borders of all
    chromosomes come from one device call (the cluster kernel: same sort, running max and border rule,
    arrops.py:459-470 with min_dist=0), the overlap check counts partners without listing pairs
    (bf_count_overlaps), and the value borders and the numbering are array operations over the whole frame."""
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    if not allow_overlaps:
        # the reference's check is `len(ops.overlap(df, df)) > len(df)` -- with the DEFAULT column names whatever
        # `cols` says (extras.py:611): the rows of that left join are the pairs plus every row without a partner
        # (null rows; not those its grouping leaves out altogether, ops._ungrouped_rows)
        checks.is_bedframe(df, raise_errors=True)
        counts = None
        if pd.api.types.is_integer_dtype(df.index.dtype):
            counts = ops._overlap_counts(df, df, None, None, None)
        if counts is None or not counts.all():
            # labels that are not integers, or rows without a partner (null rows: every other row overlaps itself):
            # what the join makes of them -- casts, blanked payload columns, rows its grouping drops -- is the
            # join's business, so it runs as in the reference; the usual frame only needs the pair count
            joined = len(ops.overlap(df, df))
        else:
            joined = int(counts.sum())
        if joined > len(df):
            raise ValueError("Not a proper bedGraph: found overlapping intervals.")

    codes, _ = pd.factorize(df[ck])  # order of first appearance, -1 for null chromosomes (dropped, as groupby does)
    rows = np.flatnonzero(codes >= 0)
    starts, ends = df[sk].to_numpy()[rows], df[ek].to_numpy()[rows]
    groups = codes[rows].astype(np.int32)
    order = np.lexsort((ends, starts, groups))  # stable, like sort_values([start, end]) inside every group
    n_groups = int(groups.max()) + 1 if len(rows) else 1
    spatial = np.empty(0, dtype=np.int64)
    if len(rows):
        res = ops._engine.cluster(groups, starts.astype(np.int64), ends.astype(np.int64), n_groups, min_dist=0,
                              want_ids=True, want_table=False)
        spatial = ops._to_host(res["ids"])[order]  # cluster of every row, in sorted order

    values = df[col].to_numpy()[rows][order]
    g = groups[order]
    border = np.ones(len(order), dtype=bool)
    if len(order) > 1:
        if values.dtype.kind == "f":
            changed = ~np.isclose(values[1:], values[:-1], equal_nan=True)
        else:
            changed = np.asarray(values[1:] != values[:-1], dtype=bool)
        border[1:] = (spatial[1:] != spatial[:-1]) | changed | (g[1:] != g[:-1])
    run = np.cumsum(border) - 1
    if reset_counter and len(order):  # numbering restarts with every chromosome
        first_of_group = np.r_[True, g[1:] != g[:-1]]
        run = run - np.maximum.accumulate(np.where(first_of_group, run, 0))
    if len(rows) == 0:
        raise ValueError("No objects to concatenate")  # pd.concat([]) of the reference's per-chromosome pieces (extras.py:651)
    out = df.iloc[rows[order]].copy()
    out[run_col] = run
    return out


def merge_runs(df, col, *, allow_overlaps=False, agg=None, cols=None):
    """Merge every run of spatially consecutive intervals sharing the value of `col` into one interval
    (extras.py:655-735); `agg` adds named aggregations over the rows of a run."""
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    marked = mark_runs(df, col, allow_overlaps=allow_overlaps, reset_counter=False, run_col="_run")
    spec = {ck: (ck, "first"), sk: (sk, "min"), ek: (ek, "max"), col: (col, "first"), **(agg or {})}
    return marked.groupby("_run").agg(**spec).reset_index(drop=True)
