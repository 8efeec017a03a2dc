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


def overlap_ordered(grp1, start1, end1, grp2, start2, end2, n_groups, device=None, max_partners=None):
    ev1, ev2, _ = overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="left")
    ev1, ev2 = sort_pairs(ev1, ev2, len(start1), len(start2))
    biggest = int(np.bincount(ev1[ev2 >= 0], minlength=1).max()) if len(ev1) else 0
    if max_partners is not None and biggest > max_partners:
        return None, None, biggest
    return ev1, ev2, biggest


def sort_pairs(ev1, ev2, max1, max2):
    a = np.where(ev1 < 0, max1 + 1, ev1)
    b = np.where(ev2 < 0, max2 + 1, ev2)
    order = np.lexsort([b, a])
    return ev1[order], ev2[order]


def count_overlaps(grp1, start1, end1, grp2, start2, end2, n_groups, closed=False, device=None):
    ev1, ev2, _ = overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=closed)
    return np.bincount(ev1, minlength=len(start1)).astype(np.int64)


# ---- prepared sets / range-sharded overlap (engine.measure_extent, prepare_set, adopt_set, overlap_prepared) ----
# An independent model of the shard-ownership rule: the whole local result from the oracle, then rows are
# dropped by classifying every pair as A-block (target starts at or after the query: arrops.py:338-339, 359-360)
# or B-block (query starts after the target: arrops.py:341-342, 365-366).
class _DoubleSet:
    def __init__(self, rows, n_groups):
        import torch

        self.n, self.n_groups = len(rows), n_groups
        self.keys = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64))  # [n, 3] = (group, start, end)
        self.ids = torch.arange(self.n, dtype=torch.int32)
        self.store = self.keys


def measure_extent(grp, start, end, n_groups, device=None):
    s, e = _i64(start), _i64(end)
    w = np.full(8 + n_groups, np.iinfo(np.int64).min, dtype=np.int64)
    w[3:8] = 0
    if len(s):
        ee = e + (e == s)
        w[0], w[1], w[2], w[3] = ~min(s.min(), ee.min()), max(s.max(), ee.max()), s.max(), (ee - s).max()
        g = np.zeros(len(s), np.int64) if grp is None else np.asarray(grp)
        np.maximum.at(w, 8 + g, np.maximum(s, ee))
    return w, False


def prepare_set(words, grp, start, end, n_groups, presorted=False, device=None, slot=None):
    s, e = _i64(start), _i64(end)
    g = np.zeros(len(s), np.int64) if grp is None else np.asarray(grp, dtype=np.int64)
    return _DoubleSet(np.stack([g, s, e], axis=1) if len(s) else np.zeros((0, 3), np.int64), n_groups)


def adopt_set(words, n, n_groups, grp=None, device=None, slot=None):
    return _DoubleSet(np.zeros((int(n), 3), np.int64), n_groups)


def set_runmax(sorted_set):
    return sorted_set


def overlap_prepared(set1, set2, how="inner", closed=False, own_range=None, row_offset1=0, row_offset2=0,
                     n_own1=None, halo_ids=None, want_group_counts=False):
    import torch

    r1, r2 = set1.keys.numpy(), set2.keys.numpy()
    g1, s1, e1 = r1[:, 0], r1[:, 1], r1[:, 2]
    g2, s2, e2 = r2[:, 0], r2[:, 1], r2[:, 2]
    nb = set1.n_groups
    ev1, ev2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how=how, closed=closed)
    grp_of = np.where(ev1 >= 0, g1[np.maximum(ev1, 0)], g2[np.maximum(ev2, 0)]) if len(ev1) else np.zeros(0, np.int64)
    paired = (ev1 >= 0) & (ev2 >= 0)
    is_a = paired & (s2[np.maximum(ev2, 0)] >= s1[np.maximum(ev1, 0)])
    is_b = paired & ~is_a
    keep = np.ones(len(ev1), dtype=bool)
    if own_range is not None:
        g_lo, s_lo, g_hi, s_hi = (int(x) for x in own_range)

        def owned(g, s):
            return ((g > g_lo) | ((g == g_lo) & (s >= s_lo))) & ((g < g_hi) | ((g == g_hi) & (s < s_hi)))

        own1 = owned(g1, s1)[np.maximum(ev1, 0)]
        own2 = owned(g2, s2)[np.maximum(ev2, 0)]
        keep = np.where(is_a, own1, np.where(is_b, own2, own1))
    ev1, ev2, grp_of, is_a, is_b = ev1[keep], ev2[keep], grp_of[keep], is_a[keep], is_b[keep]
    counts = np.zeros((nb, 4), dtype=np.int64)
    np.add.at(counts[:, 0], grp_of[is_a], 1)
    np.add.at(counts[:, 1], grp_of[is_b], 1)
    np.add.at(counts[:, 2], grp_of[(ev2 < 0)], 1)
    n_own = set1.n if n_own1 is None else int(n_own1)
    out1 = ev1 + row_offset1
    if n_own < set1.n:
        halo = np.asarray(halo_ids, dtype=np.int64)
        out1 = np.where(ev1 >= n_own, halo[np.maximum(ev1 - n_own, 0)], out1)
    res = (torch.from_numpy(out1), torch.from_numpy(np.where(ev2 >= 0, ev2 + row_offset2, -1)), int(paired[keep].sum()))
    return (*res, counts) if want_group_counts else res


def codes_present(codes, n_categories, device=None):
    codes = np.asarray(codes)
    flags = np.zeros(int(n_categories) + 1, dtype=bool)
    flags[np.unique(codes).astype(np.int64) + 1] = True
    return codes, flags


def codes_apply(codes_dev, lut, device=None):
    return np.asarray(lut, dtype=np.int32)[np.asarray(codes_dev).astype(np.int64) + 1]
