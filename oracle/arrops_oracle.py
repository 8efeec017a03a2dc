"""CPU ORACLE (test infrastructure, NOT the product path).

A numpy restatement of the array-level interval algorithms of the reference
(`/root/reference/src/bioframe/core/arrops.py`, bioframe v0.8.0).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
legs may import this module; the product (`bioframe_b200`) never does.

Parity status: PINNED.  Every function is checked against outputs of the
reference itself (imported in the build container by
`tests/golden/make_golden.py`, results committed as `tests/golden/*.npz`) and
against the hand-computed vectors of SURVEY.md §8c.

Each function states the reference lines it follows.  One deliberate
definition: where the reference calls an *unstable* `np.argsort`
(arrops.py:562,580) the oracle uses a stable sort; the two agree on tie-free
keys, which is what the golden vectors use.
"""
from __future__ import annotations

import numpy as np

INT64_MAX = np.iinfo(np.int64).max
_EMPTY = np.empty(0, dtype=np.int64)


def pseudo_ends(starts, ends):
    """Points [x,x) behave as [x,x+1) inside overlap (arrops.py:271-287)."""
    out = np.array(ends, copy=True)
    out[ends == starts] += 1
    return out


def expand_ranges(lo, hi):
    """Concatenation of arange(lo[k], hi[k]) for all k (arrops.py:80-132)."""
    lo = np.asarray(lo)
    hi = np.asarray(hi)
    width = hi - lo
    total = int(width.sum())
    first_out = np.cumsum(width) - width  # output slot of each range's first element
    return np.arange(total) - np.repeat(first_out - lo, width)


def overlap_intervals(starts1, ends1, starts2, ends2, closed=False, sort=False):
    """All (i, j) with s1<e2' and s2<e1' in the reference's A-block/B-block
    order (arrops.py:290-375)."""
    s1 = np.asarray(starts1)
    s2 = np.asarray(starts2)
    e1 = pseudo_ends(s1, np.asarray(ends1))
    e2 = pseudo_ends(s2, np.asarray(ends2))

    # stable (start, end', row) order of each set -- arrops.py:332-335
    perm1 = np.lexsort([e1, s1])
    perm2 = np.lexsort([e2, s2])
    S1, E1 = s1[perm1], e1[perm1]
    S2, E2 = s2[perm2], e2[perm2]

    hi_side = "right" if closed else "left"
    # A-block: per sorted row of set 1, targets whose start lies in [s1, e1')
    a_lo = np.searchsorted(S2, S1, "left")  # arrops.py:338
    a_hi = np.searchsorted(S2, E1, hi_side)  # arrops.py:339
    # B-block: per sorted row of set 2, queries whose start lies in (s2, e2')
    b_lo = np.searchsorted(S1, S2, "right")  # arrops.py:341 ("right": no double report)
    b_hi = np.searchsorted(S1, E2, hi_side)  # arrops.py:342

    a_keep = a_hi > a_lo  # arrops.py:345-354
    b_keep = b_hi > b_lo
    a_lo, a_hi = a_lo[a_keep], a_hi[a_keep]
    b_lo, b_hi = b_lo[b_keep], b_hi[b_keep]

    # arrops.py:357-368
    ids1 = np.concatenate(
        [np.repeat(perm1[a_keep], a_hi - a_lo), perm1[expand_ranges(b_lo, b_hi)]]
    )
    ids2 = np.concatenate(
        [perm2[expand_ranges(a_lo, a_hi)], np.repeat(perm2[b_keep], b_hi - b_lo)]
    )
    if sort:  # arrops.py:370-373
        o = np.lexsort([ids2, ids1])
        ids1, ids2 = ids1[o], ids2[o]
    return ids1, ids2


def overlap_intervals_outer(starts1, ends1, starts2, ends2, closed=False):
    """Pairs plus ascending unpaired ids of both sets (arrops.py:378-412)."""
    n1, n2 = len(starts1), len(starts2)
    i1, i2 = overlap_intervals(starts1, ends1, starts2, ends2, closed=closed)
    u1 = np.setdiff1d(np.arange(n1), i1) if n1 > 0 else np.array([], dtype=int)
    u2 = np.setdiff1d(np.arange(n2), i2) if n2 > 0 else np.array([], dtype=int)
    return i1, i2, u1, u2


def merge_intervals(starts, ends, min_dist=0):
    """Cluster ids in input order + merged spans (arrops.py:415-479; App. A3).
    No point adjustment here: raw ends."""
    s = np.asarray(starts)
    e = np.asarray(ends)
    n = s.shape[0]
    perm = np.lexsort([e, s])  # arrops.py:459
    S = s[perm]
    R = np.maximum.accumulate(e[perm])  # arrops.py:462 running max of end

    is_head = np.ones(n + 1, dtype=bool)  # arrops.py:463-465 (both sentinels True)
    if min_dist is not None:
        is_head[1:-1] = S[1:] > R[:-1] + min_dist  # arrops.py:468
    else:
        is_head[1:-1] = S[1:] >= R[:-1]  # arrops.py:470

    ids_sorted = np.cumsum(is_head)[:-1] - 1  # arrops.py:472
    ids = np.full(n, -1)
    ids[perm] = ids_sorted  # arrops.py:473-474
    return ids, S[is_head[:-1]], R[is_head[1:]]  # arrops.py:476-477


def complement_intervals(starts, ends, bounds=(0, INT64_MAX)):
    """Gaps between merged intervals inside `bounds` (arrops.py:482-503; A4)."""
    _, ms, me = merge_intervals(starts, ends, min_dist=0)
    lo = np.searchsorted(me, bounds[0], "right")  # arrops.py:489
    hi = np.searchsorted(ms, bounds[1], "left")  # arrops.py:490
    ms, me = ms[lo:hi], me[lo:hi]
    gs = np.r_[bounds[0], me]  # arrops.py:496
    ge = np.r_[ms, bounds[1]]  # arrops.py:497
    first = 1 if gs[0] >= ge[0] else 0  # arrops.py:498
    last = -1 if gs[-1] >= ge[-1] else None  # arrops.py:499
    return gs[first:last], ge[first:last]


def _neighbours_one_side(starts1, ends1, starts2, ends2, direction, k):
    """k nearest non-overlapping targets on one side (arrops.py:506-598).
    Stable sort of the targets replaces the reference's unstable argsort."""
    s1, e1 = np.asarray(starts1), np.asarray(ends1)
    s2, e2 = np.asarray(starts2), np.asarray(ends2)
    n1, n2 = s1.shape[0], s2.shape[0]
    if k <= 0:
        return np.array([], dtype=int), np.array([], dtype=int)
    if direction == "left":
        by_end = np.argsort(e2, kind="stable")  # arrops.py:562
        stop = np.searchsorted(e2[by_end], s1, "right")  # arrops.py:569: #(e2 <= s1)
        begin = np.maximum(stop - k, 0)  # arrops.py:570
        order = by_end
    elif direction == "right":
        by_start = np.argsort(s2, kind="stable")  # arrops.py:580
        begin = np.searchsorted(s2[by_start], e1, "left")  # arrops.py:587: #(s2 < e1)
        stop = np.minimum(begin + k, n2)  # arrops.py:588
        order = by_start
    else:
        raise ValueError(direction)
    q = np.repeat(np.arange(n1), stop - begin)
    t = order[expand_ranges(begin, stop)] if q.size else np.array([], dtype=int)
    return q, t


def closest_intervals(
    starts1,
    ends1,
    starts2=None,
    ends2=None,
    k=1,
    tie_arr=None,
    ignore_overlaps=False,
    ignore_upstream=False,
    ignore_downstream=False,
    along=None,
):
    """k closest targets per query (arrops.py:601-754).
    `tie_arr` is not supported (broken upstream)."""
    if tie_arr is not None:
        raise NotImplementedError("tie_arr: reference path is broken (arrops.py:740)")
    s1, e1 = np.asarray(starts1), np.asarray(ends1)
    # arrops.py:650-659 -- overlap candidates (distance 0)
    if ignore_overlaps:
        o1 = o2 = np.array([], dtype=int)
        if starts2 is None and ends2 is None:
            raise IndexError("reference quirk B5: self-closest with ignore_overlaps")
        s2, e2 = np.asarray(starts2), np.asarray(ends2)
    elif starts2 is None and ends2 is None:
        s2, e2 = s1, e1
        o1, o2 = overlap_intervals(s1, e1, s2, e2)
        keep = o1 != o2
        o1, o2 = o1[keep], o2[keep]
    else:
        s2, e2 = np.asarray(starts2), np.asarray(ends2)
        o1, o2 = overlap_intervals(s1, e1, s2, e2)

    n = s1.shape[0]
    if along is None:
        along = np.ones(n, dtype=bool)
    along = np.asarray(along, dtype=bool)
    k_up = 0 if ignore_upstream else k
    k_dn = 0 if ignore_downstream else k

    # arrops.py:666-720: "+" rows look left for upstream, "-" rows look right.
    fwd = np.flatnonzero(along)
    rev = np.flatnonzero(~along)
    fu1, fu2 = _neighbours_one_side(s1[fwd], e1[fwd], s2, e2, "left", k_up)
    fd1, fd2 = _neighbours_one_side(s1[fwd], e1[fwd], s2, e2, "right", k_dn)
    ru1, ru2 = _neighbours_one_side(s1[rev], e1[rev], s2, e2, "right", k_up)
    rd1, rd2 = _neighbours_one_side(s1[rev], e1[rev], s2, e2, "left", k_dn)
    l1 = np.concatenate([fwd[fu1], rev[rd1]]).astype(np.int64)
    l2 = np.concatenate([fu2, rd2]).astype(np.int64)
    r1 = np.concatenate([rev[ru1], fwd[fd1]]).astype(np.int64)
    r2 = np.concatenate([ru2, fd2]).astype(np.int64)

    # arrops.py:724-730: +1 so true overlaps (0) sort before abutting neighbours
    dl = s1[l1] - e2[l2] + 1
    dr = s2[r2] - e1[r1] + 1
    q = np.concatenate([l1, r1, o1])
    t = np.concatenate([l2, r2, o2])
    d = np.concatenate([dl, dr, np.zeros(o1.shape[0])])
    if q.shape[0] == 0:
        return np.array([], dtype=int), np.array([], dtype=int)

    o = np.lexsort([t, d, q])  # arrops.py:738
    q, t = q[o], t[o]
    edges = np.flatnonzero(np.r_[True, q[:-1] != q[1:], True])  # arrops.py:746
    run_lo = edges[:-1]
    run_len = np.minimum(k, edges[1:] - run_lo)  # arrops.py:749-752
    pick = expand_ranges(run_lo, run_lo + run_len)
    return q[pick], t[pick]
