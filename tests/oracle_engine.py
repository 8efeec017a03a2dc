"""TEST DOUBLE for `bioframe_b200.engine`, backed by the CPU oracle.

The pandas glue of `bioframe_b200.ops` / `core.arrops` cannot reach a GPU in
the build container; these stand-ins have the engine's signatures and return
numpy arrays, so the `-m "not gpu"` tests can check the glue (group encoding,
NA handling, frame assembly, errors) against the reference.  On the GPU box
the same frame-level tests run with the real engine.  Never imported by the
product."""
import numpy as np

from oracle import ops_oracle as oo


def _i64(x):
    return np.asarray(x, dtype=np.int64) if x is not None else None


def overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=False, device=None,
                    row_offset1=0, row_offset2=0):
    s1, e1, s2, e2 = _i64(start1), _i64(end1), _i64(start2), _i64(end2)
    ev1, ev2 = oo.overlap_intidxs_coded(grp1, s1, e1, grp2, s2, e2, n_groups, how=how, closed=closed)
    n_pairs = int(((ev1 >= 0) & (ev2 >= 0)).sum())
    ev1 = np.where(ev1 >= 0, ev1 + row_offset1, ev1)
    ev2 = np.where(ev2 >= 0, ev2 + row_offset2, ev2)
    return ev1, ev2, n_pairs


def closest_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, k=1, ignore_overlaps=False,
                    ignore_upstream=False, ignore_downstream=False, along=None, self_closest=False, device=None):
    ev1, ev2 = oo.closest_intidxs_coded(
        grp1, _i64(start1), _i64(end1), grp2, _i64(start2), _i64(end2), n_groups, k=k,
        ignore_overlaps=ignore_overlaps, ignore_upstream=ignore_upstream, ignore_downstream=ignore_downstream,
        along=along, self_closest=self_closest,
    )
    return ev1, ev2, int((ev2 >= 0).sum())


def cluster(grp, start, end, n_groups, min_dist=0, want_ids=True, want_table=True, want_row_spans=False, device=None):
    s, e = _i64(start), _i64(end)
    ids, cs, ce, cn = oo.cluster_coded(grp, s, e, n_groups, min_dist)
    out = dict(ids=ids if (want_ids or want_row_spans) else None, n_clusters=len(cs), cgrp=None, cstart=None,
               cend=None, ccount=None, row_start=None, row_end=None)
    if want_table:
        out.update(cgrp=oo.cluster_coded.last_groups.astype(np.int32), cstart=cs, cend=ce, ccount=cn)
    if want_row_spans:
        out.update(row_start=cs[ids], row_end=ce[ids])
    return out


def coverage(grp1, start1, end1, grp2, start2, end2, n_groups, device=None):
    return oo.coverage_coded(grp1, _i64(start1), _i64(end1), grp2, _i64(start2), _i64(end2), n_groups)


def download_rows(t, out=None, n_threads=None):
    t = np.asarray(t, dtype=np.int64)
    if out is None:
        return t.copy()
    out[: len(t)] = t
    return out[: len(t)]


def to_device(x, dtype="int64", device=None):
    return np.asarray(x, dtype=dtype)


def gather_rows(column, idx, want_valid=False, device=None):
    column, idx = np.asarray(column), _i64(idx)
    out = np.where(idx >= 0, column[np.maximum(idx, 0)], column.dtype.type(0)) if len(column) else np.zeros(len(idx), column.dtype)
    out = out.astype(column.dtype, copy=False)
    return (out, (idx >= 0).astype(np.uint8)) if want_valid else out


def pair_coords(ev1, ev2, start1, end1, start2, end2, want_overlap=True, want_distance=False, device=None):
    a, b = _i64(ev1), _i64(ev2)
    s1, e1, s2, e2 = _i64(start1), _i64(end1), _i64(start2), _i64(end2)
    ok = (a >= 0) & (b >= 0)
    if len(s1) == 0 or len(s2) == 0:
        z = np.zeros(len(a), np.int64)
        return (z, z.copy(), z.copy() if want_distance else None) if want_overlap else (None, None, z)
    os_ = np.where(ok, np.maximum(s1[a], s2[b]), 0)
    oe = np.where(ok, np.minimum(e1[a], e2[b]), 0)
    dist = np.where(ok, np.maximum(0, np.maximum(s1[a] - e2[b], s2[b] - e1[a])), 0)
    return (os_ if want_overlap else None, oe if want_overlap else None, dist if want_distance else None)


def subtract(grp1, start1, end1, grp2, start2, end2, n_groups, want_sub_index=True, device=None):
    row, ps, pe, sub = oo.subtract_coded(grp1, _i64(start1), _i64(end1), grp2, _i64(start2), _i64(end2), n_groups)
    return row, ps, pe, (sub if want_sub_index else None)


def complement(region, start, end, bounds_lo, bounds_hi, device=None):
    b0, b1 = _i64(bounds_lo), _i64(bounds_hi)
    if len(b0) == 0:
        z = np.empty(0, np.int64)
        return z.astype(np.int32), z, z
    region = np.zeros(len(start), np.int32) if region is None else np.asarray(region)
    oc, os_, oe = oo.complement_core(region, _i64(start), _i64(end), list(zip(b0.tolist(), b1.tolist())))
    return oc.astype(np.int32), os_, oe


def overlap_ordered(grp1, start1, end1, grp2, start2, end2, n_groups, device=None):
    ev1, ev2, _ = overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="left")
    ev1, ev2 = sort_pairs(ev1, ev2, len(start1), len(start2))
    biggest = int(np.bincount(ev1[ev2 >= 0], minlength=1).max()) if len(ev1) else 0
    return ev1, ev2, biggest


def sort_pairs(ev1, ev2, max1, max2):
    a = np.where(ev1 < 0, max1 + 1, ev1)
    b = np.where(ev2 < 0, max2 + 1, ev2)
    order = np.lexsort([b, a])
    return ev1[order], ev2[order]


def count_overlaps(grp1, start1, end1, grp2, start2, end2, n_groups, closed=False, device=None):
    ev1, ev2, _ = overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=closed)
    return np.bincount(ev1, minlength=len(start1)).astype(np.int64)
