"""CPU ORACLE (test infrastructure, NOT the product path) -- frame level.

Restates the per-group loops of the reference's `src/bioframe/ops.py`
(v0.8.0) on top of `oracle.arrops_oracle`: `_overlap_intidxs`
(ops.py:242-338), `_closest_intidxs` (ops.py:919-1040), the group loops of
`cluster`/`merge` (ops.py:637-674 / 776-810), the region loop of `complement`
(ops.py:1644-1685) and `coverage` (ops.py:842-916).  Only tests, smoke() and
bench.py's CPU legs import this.  Parity status: PINNED by
`tests/golden/*.npz` (reference outputs generated in the build container).

The chromosome-block order of overlap/closest is the iteration order of a
Python `set` in the reference (ops.py:289,985), i.e. PYTHONHASHSEED-dependent;
`group_order=` lets a test replay the order a golden vector was recorded with.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import arrops_oracle as ao

_COLS = ("chrom", "start", "end")


def _group_rows(df, by, dropna):
    return df.groupby(by, observed=True, dropna=dropna).indices


def _set_order(g1, g2):
    return list(set.union(set(g1), set(g2)))  # ops.py:289


def overlap_intidxs(df1, df2, how="left", cols1=None, cols2=None, on=None, group_order=None):
    """(events1, events2) global row positions, -1 = no partner (ops.py:242-338)."""
    ck1, sk1, ek1 = cols1 or _COLS
    ck2, sk2, ek2 = cols2 or _COLS
    by1 = [ck1] + (list(on) if on else [])
    by2 = [ck2] + (list(on) if on else [])
    g1 = _group_rows(df1, by1, dropna=False)  # ops.py:287
    g2 = _group_rows(df2, by2, dropna=False)
    keys = _set_order(g1, g2) if group_order is None else group_order
    s1, e1 = df1[sk1].values, df1[ek1].values
    s2, e2 = df2[sk2].values, df2[ek2].values
    take_left = how in ("left", "outer")
    take_right = how in ("right", "outer")
    none = np.array([], dtype=np.int64)

    out1, out2 = [], []
    for key in keys:  # ops.py:301
        r1 = g1.get(key, none)
        r2 = g2.get(key, none)
        if len(r1) and len(r2):
            i, j = ao.overlap_intervals(s1[r1], e1[r1], s2[r2], e2[r2])  # ops.py:311
            out1.append(r1[i])
            out2.append(r2[j])
            if take_left:  # ops.py:228-239 (_minus) + 321-323
                lone = np.setdiff1d(np.arange(len(r1)), i)
                if len(lone):
                    out1.append(r1[lone])
                    out2.append(np.full(len(lone), -1))
            if take_right:  # ops.py:324-326
                lone = np.setdiff1d(np.arange(len(r2)), j)
                if len(lone):
                    out1.append(np.full(len(lone), -1))
                    out2.append(r2[lone])
        elif len(r1) and take_left:  # ops.py:327-330
            out1.append(r1)
            out2.append(np.full(len(r1), -1))
        elif len(r2) and take_right:  # ops.py:331-334
            out1.append(np.full(len(r2), -1))
            out2.append(r2)
    if not out1:
        return none.copy(), none.copy()
    return np.concatenate(out1).astype(np.int64), np.concatenate(out2).astype(np.int64)


def overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, n_groups, how="inner", closed=False):
    """`overlap_intidxs` on group-coded arrays with the groups visited in code
    order -- the same contract as the C ABI's bf_overlap_* (ops.py:301-338)."""
    g1, g2 = np.asarray(g1), np.asarray(g2)
    take_left = how in ("left", "outer")
    take_right = how in ("right", "outer")
    out1, out2 = [], []
    if n_groups == 1:
        rows1 = [np.arange(len(s1))]
        rows2 = [np.arange(len(s2))]
    else:
        o1 = np.argsort(g1, kind="stable")
        o2 = np.argsort(g2, kind="stable")
        b1 = np.searchsorted(g1[o1], np.arange(n_groups + 1))
        b2 = np.searchsorted(g2[o2], np.arange(n_groups + 1))
        rows1 = [o1[b1[c] : b1[c + 1]] for c in range(n_groups)]
        rows2 = [o2[b2[c] : b2[c + 1]] for c in range(n_groups)]
    for r1, r2 in zip(rows1, rows2):
        if len(r1) and len(r2):
            i, j = ao.overlap_intervals(s1[r1], e1[r1], s2[r2], e2[r2], closed=closed)
            out1.append(r1[i])
            out2.append(r2[j])
            if take_left:
                lone = np.setdiff1d(np.arange(len(r1)), i)
                out1.append(r1[lone])
                out2.append(np.full(len(lone), -1))
            if take_right:
                lone = np.setdiff1d(np.arange(len(r2)), j)
                out1.append(np.full(len(lone), -1))
                out2.append(r2[lone])
        elif len(r1) and take_left:
            out1.append(r1)
            out2.append(np.full(len(r1), -1))
        elif len(r2) and take_right:
            out1.append(np.full(len(r2), -1))
            out2.append(r2)
    if not out1:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    return np.concatenate(out1).astype(np.int64), np.concatenate(out2).astype(np.int64)


def closest_intidxs(
    df1,
    df2=None,
    k=1,
    ignore_overlaps=False,
    ignore_upstream=False,
    ignore_downstream=False,
    direction_col=None,
    cols1=None,
    cols2=None,
    group_order=None,
):
    """(events1, events2); unmatched queries get -1 (ops.py:919-1040)."""
    ck1, sk1, ek1 = cols1 or _COLS
    ck2, sk2, ek2 = cols2 or _COLS
    self_mode = df2 is None or df2 is df1
    if self_mode:
        if len(df1) == 1:
            raise ValueError(
                "df1 must have more than one interval to find closest non-identical interval"
            )
        df2 = df1
    g1 = df1.groupby(ck1, observed=True).indices  # ops.py:983 (dropna default: NA rows dropped)
    g2 = df2.groupby(ck2, observed=True).indices
    keys = _set_order(g1, g2) if group_order is None else group_order
    s1, e1 = df1[sk1].values, df1[ek1].values
    s2, e2 = df2[sk2].values, df2[ek2].values
    strand = df1[direction_col].values if direction_col is not None else None
    none = np.array([], dtype=np.int64)

    out1, out2 = [], []
    for key in keys:  # ops.py:997
        r1 = g1.get(key, none)
        r2 = g2.get(key, none)
        if len(r1) and len(r2):
            along = strand[r1] != "-" if strand is not None else np.ones(len(r1), dtype=bool)
            i, j = ao.closest_intervals(
                s1[r1],
                e1[r1],
                None if self_mode else s2[r2],
                None if self_mode else e2[r2],
                k=k,
                ignore_overlaps=ignore_overlaps,
                ignore_upstream=ignore_upstream,
                ignore_downstream=ignore_downstream,
                along=along,
            )
            lone = np.setdiff1d(np.arange(len(r1)), i)  # ops.py:1027
            out1 += [r1[i], r1[lone]]
            out2 += [r2[j], np.full(len(lone), -1)]
        elif len(r1):  # ops.py:1033-1035
            out1.append(r1)
            out2.append(np.full(len(r1), -1))
    if not out1:
        return none.copy(), none.copy()
    return np.concatenate(out1).astype(np.int64), np.concatenate(out2).astype(np.int64)


def cluster_core(df, min_dist=0, cols=None, on=None):
    """cluster ids per row (NA rows excluded: -1) and the cluster table
    (group key, start, end, n_intervals) in the reference's order: groups in
    sorted key order, ids offset by clusters so far (ops.py:631-674)."""
    ck, sk, ek = cols or _COLS
    by = [ck] + (list(on) if on else [])
    df = df.reset_index(drop=True)
    groups = df.groupby(by, observed=True).indices  # sorted keys, NA keys dropped
    ids = np.full(len(df), -1, dtype=np.int64)
    keys, cs, ce, cn = [], [], [], []
    base = 0
    for key, rows in groups.items():
        if len(rows) == 0:
            continue
        gi, gs, ge = ao.merge_intervals(
            df[sk].values[rows].astype(np.int64), df[ek].values[rows].astype(np.int64), min_dist
        )  # ops.py:647-651
        ids[rows] = gi + base  # ops.py:654-658
        base += len(gs)
        keys += [key] * len(gs)
        cs.append(gs)
        ce.append(ge)
        cn.append(np.bincount(gi))  # ops.py:653
    cat = lambda xs: np.concatenate(xs) if xs else np.empty(0, dtype=np.int64)  # noqa: E731
    return ids, keys, cat(cs), cat(ce), cat(cn)


def coverage_core(df1, df2, cols1=None, cols2=None):
    """Per-row covered base pairs of df1 by merged df2 (ops.py:842-916),
    computed the reference's way: merge, left overlap, clip, sum."""
    ck1, sk1, ek1 = cols1 or _COLS
    ck2, sk2, ek2 = cols2 or _COLS
    d2 = df2[[ck2, sk2, ek2]].dropna()
    _, keys, ms, me, _ = cluster_core(d2, 0, cols=(ck2, sk2, ek2))
    merged = pd.DataFrame({ck2: pd.Series(keys, dtype=object), sk2: ms, ek2: me})
    d1 = df1.reset_index(drop=True)
    i, j = overlap_intidxs(d1, merged, how="inner", cols1=cols1, cols2=(ck2, sk2, ek2))
    s1, e1 = d1[sk1].values, d1[ek1].values
    ov = np.minimum(e1[i], me[j]) - np.maximum(s1[i], ms[j])
    out = np.zeros(len(d1), dtype=np.int64)
    np.add.at(out, i, ov.astype(np.int64))
    return out


def complement_core(region_codes, starts, ends, region_bounds):
    """Per view region (ascending code order, = sorted region names in
    ops.py:1645-1648) the gaps of the clipped intervals inside the region's
    bounds (ops.py:1665-1669).  Regions with no interval yield the whole
    region (ops.py:1652-1657).  -> (codes, gap_starts, gap_ends)."""
    region_codes = np.asarray(region_codes)
    oc, os_, oe = [], [], []
    for code, (b0, b1) in enumerate(region_bounds):
        rows = np.flatnonzero(region_codes == code)
        if len(rows) == 0:
            gs, ge = np.array([b0]), np.array([b1])
        else:
            gs, ge = ao.complement_intervals(starts[rows], ends[rows], bounds=(b0, b1))
        oc.append(np.full(len(gs), code, dtype=np.int64))
        os_.append(np.asarray(gs, dtype=np.int64))
        oe.append(np.asarray(ge, dtype=np.int64))
    return np.concatenate(oc), np.concatenate(os_), np.concatenate(oe)


def _rows_by_code(g, n, n_groups):
    if n_groups == 1 or g is None:
        return [np.arange(n)]
    g = np.asarray(g)
    o = np.argsort(g, kind="stable")
    b = np.searchsorted(g[o], np.arange(n_groups + 1))
    return [o[b[c] : b[c + 1]] for c in range(n_groups)]


def cluster_coded(g, s, e, n_groups, min_dist=0):
    """cluster ids per row + cluster table (start, end, n_intervals) with the
    groups visited in code order -- the contract of the C ABI's bf_cluster_*
    (ops.py:637-674 with group codes standing in for sorted group keys)."""
    s, e = np.asarray(s, dtype=np.int64), np.asarray(e, dtype=np.int64)
    ids = np.full(len(s), -1, dtype=np.int64)
    cs, ce, cn, cg = [], [], [], []
    base = 0
    for c, rows in enumerate(_rows_by_code(g, len(s), n_groups)):
        if len(rows) == 0:
            continue
        gi, gs, ge = ao.merge_intervals(s[rows], e[rows], min_dist)
        ids[rows] = gi + base
        base += len(gs)
        cs.append(gs)
        ce.append(ge)
        cn.append(np.bincount(gi))
        cg.append(np.full(len(gs), c))
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.empty(0, dtype=np.int64)  # noqa: E731
    cluster_coded.last_groups = cat(cg)
    return ids, cat(cs), cat(ce), cat(cn)


def coverage_coded(g1, s1, e1, g2, s2, e2, n_groups):
    """coverage on group-coded arrays, the reference's way: merge set 2, inner
    overlap, clip, per-query sum (ops.py:888-912)."""
    s1, e1 = np.asarray(s1, dtype=np.int64), np.asarray(e1, dtype=np.int64)
    _, ms, me, _ = cluster_coded(g2, s2, e2, n_groups, 0)
    mg = cluster_coded.last_groups.astype(np.int32)
    i, j = overlap_intidxs_coded(g1, s1, e1, mg, ms, me, n_groups, how="inner")
    ov = np.minimum(e1[i], me[j]) - np.maximum(s1[i], ms[j])
    out = np.zeros(len(s1), dtype=np.int64)
    np.add.at(out, i, ov)
    return out


def subtract_coded(g1, s1, e1, g2, s2, e2, n_groups):
    """ops.subtract (ops.py:1243-1330) on group-coded arrays, composed the reference's way: per group the
    targets are clipped to the view [0, INT64_MAX) (inner overlap with the view region, ops.py:1614-1622),
    complemented (arrops.py:482-503; a group without targets yields the whole view, ops.py:1652-1657), the
    queries are overlapped with the gaps (arrops.py:290-375) and clipped to them (ops.py:484-492); rows are
    ordered by (query row, gap) as keep_order=True does (ops.py:549-550) and the gap ordinals are rebased per
    query (ops.py:1324-1326).  -> (row, start, end, sub_index)."""
    s1, e1 = np.asarray(s1, dtype=np.int64), np.asarray(e1, dtype=np.int64)
    s2, e2 = np.asarray(s2, dtype=np.int64), np.asarray(e2, dtype=np.int64)
    rows1 = _rows_by_code(g1, len(s1), n_groups)
    rows2 = _rows_by_code(g2, len(s2), n_groups)
    out = [[], [], [], []]
    view_s, view_e = np.array([0], dtype=np.int64), np.array([ao.INT64_MAX], dtype=np.int64)
    for r1, r2 in zip(rows1, rows2):
        if len(r1) == 0:
            continue
        _, inside = ao.overlap_intervals(view_s, view_e, s2[r2], e2[r2])
        if len(inside):
            cs = np.maximum(s2[r2][inside], 0)
            ce = np.minimum(e2[r2][inside], ao.INT64_MAX)
            gs, ge = ao.complement_intervals(cs, ce, bounds=(0, ao.INT64_MAX))
            gs, ge = np.asarray(gs, dtype=np.int64), np.asarray(ge, dtype=np.int64)
        else:
            gs, ge = view_s, view_e
        i, j = ao.overlap_intervals(s1[r1], e1[r1], gs, ge)
        order = np.lexsort([j, r1[i]])
        i, j = i[order], j[order]
        row = r1[i]
        out[0].append(row)
        out[1].append(np.maximum(s1[row], gs[j]))
        out[2].append(np.minimum(e1[row], ge[j]))
        first = np.r_[True, row[1:] != row[:-1]] if len(row) else np.empty(0, dtype=bool)
        start_of_run = np.maximum.accumulate(np.where(first, np.arange(len(row)), 0)) if len(row) else np.empty(0, dtype=np.int64)
        out[3].append(j - j[start_of_run] if len(row) else j)
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.empty(0, dtype=np.int64)  # noqa: E731
    row, ps, pe, sub = (cat(x) for x in out)
    order = np.argsort(row, kind="stable")  # groups interleave in row order
    return row[order], ps[order], pe[order], sub[order]


def closest_intidxs_coded(g1, s1, e1, g2, s2, e2, n_groups, k=1, ignore_overlaps=False, ignore_upstream=False,
                          ignore_downstream=False, along=None, self_closest=False):
    """`closest_intidxs` on group-coded arrays, groups visited in code order --
    the contract of the C ABI's bf_closest_* (ops.py:997-1035)."""
    s1, e1 = np.asarray(s1, dtype=np.int64), np.asarray(e1, dtype=np.int64)
    if self_closest:
        g2, s2, e2 = g1, s1, e1
    s2, e2 = np.asarray(s2, dtype=np.int64), np.asarray(e2, dtype=np.int64)
    rows1 = _rows_by_code(g1, len(s1), n_groups)
    rows2 = _rows_by_code(g2, len(s2), n_groups)
    out1, out2 = [], []
    for r1, r2 in zip(rows1, rows2):
        if len(r1) and len(r2):
            al = np.asarray(along, dtype=bool)[r1] if along is not None else None
            i, j = ao.closest_intervals(
                s1[r1], e1[r1], None if self_closest else s2[r2], None if self_closest else e2[r2], k=k,
                ignore_overlaps=ignore_overlaps, ignore_upstream=ignore_upstream,
                ignore_downstream=ignore_downstream, along=al,
            )
            lone = np.setdiff1d(np.arange(len(r1)), i)
            out1 += [r1[i], r1[lone]]
            out2 += [r2[j], np.full(len(lone), -1)]
        elif len(r1):
            out1.append(r1)
            out2.append(np.full(len(r1), -1))
    if not out1:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    return np.concatenate(out1).astype(np.int64), np.concatenate(out2).astype(np.int64)
