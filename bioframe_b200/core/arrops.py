"""Array-level API: the host-side mirror of `bioframe.core.arrops` (reference
src/bioframe/core/arrops.py, v0.8.0) for the functions on the hot path.  Same
signatures and results (int64 numpy arrays in the reference's order); each
call is ONE group handed to the CUDA engine.  No CPU fallback.

  overlap_intervals        arrops.py:290-375
  overlap_intervals_outer  arrops.py:378-412
  merge_intervals          arrops.py:415-479
  complement_intervals     arrops.py:482-503
  closest_intervals        arrops.py:601-754
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from .. import engine as _engine

INT64_MAX = np.iinfo(np.int64).max

__all__ = [
    "overlap_intervals",
    "overlap_intervals_outer",
    "merge_intervals",
    "complement_intervals",
    "closest_intervals",
]


def _arrays(*vecs):
    """int64 numpy views of the inputs; a pandas Series warns like the reference
    does (arrops.py:308-315): its index is ignored."""
    for vec in vecs:
        if isinstance(vec, pd.Series):
            warnings.warn(
                "One of the inputs is provided as pandas.Series and its index will be ignored.",
                SyntaxWarning,
                stacklevel=3,
            )
    out = []
    for vec in vecs:
        arr = np.asarray(vec)
        if arr.dtype.kind == "f":
            # the kernels compute in int64: whole-number floats pass (results come back in the input dtype, as the
            # reference's do), anything else is refused rather than truncated (the reference would carry NaN /
            # fractions through numpy)
            if not (np.isfinite(arr).all() and (arr == np.floor(arr)).all()):
                raise ValueError("coordinates must be whole numbers (NaN / fractional values are not supported)")
        out.append(np.ascontiguousarray(arr, dtype=np.int64))
    return out


def _dtype_of(vec):
    return np.asarray(vec).dtype if not isinstance(vec, np.ndarray) else vec.dtype


def _host(t):
    return t.cpu().numpy()


def overlap_intervals(starts1, ends1, starts2, ends2, closed=False, sort=False):
    """Indices (ids1, ids2) of all overlapping pairs.  Points [x, x) count as
    [x, x+1); `closed=True` treats touching intervals as overlapping;
    `sort=True` orders the pairs by (id1, id2)."""
    s1, e1, s2, e2 = _arrays(starts1, ends1, starts2, ends2)
    ev1, ev2, _ = _engine.overlap_intidxs(None, s1, e1, None, s2, e2, 1, how="inner", closed=closed)
    ids1, ids2 = _host(ev1), _host(ev2)
    if sort:
        order = np.lexsort([ids2, ids1])
        ids1, ids2 = ids1[order], ids2[order]
    return ids1, ids2


def overlap_intervals_outer(starts1, ends1, starts2, ends2, closed=False):
    """Overlapping pairs plus the ascending ids of the intervals of either set
    that overlap nothing: (ids1, ids2, no_overlap_ids1, no_overlap_ids2)."""
    s1, e1, s2, e2 = _arrays(starts1, ends1, starts2, ends2)
    ev1, ev2, n_pairs = _engine.overlap_intidxs(None, s1, e1, None, s2, e2, 1, how="outer", closed=closed)
    ev1, ev2 = _host(ev1), _host(ev2)
    lone1 = ev1[n_pairs:][ev2[n_pairs:] == -1]
    lone2 = ev2[n_pairs:][ev1[n_pairs:] == -1]
    return ev1[:n_pairs], ev2[:n_pairs], lone1, lone2


def merge_intervals(starts, ends, min_dist=0):
    """(cluster_ids, cluster_starts, cluster_ends): ids in input order;
    `min_dist=None` keeps adjacent intervals apart, a number merges intervals
    separated by that distance or less."""
    s, e = _arrays(starts, ends)
    md = None if min_dist is None else int(np.floor(min_dist))
    if md is not None and md < 0:
        raise ValueError("min_dist>=0 currently required")
    res = _engine.cluster(None, s, e, 1, min_dist=md)
    # the spans are elements of the inputs (arrops.py:472-473): they keep the inputs' dtypes
    return (_host(res["ids"]), _host(res["cstart"]).astype(_dtype_of(starts), copy=False),
            _host(res["cend"]).astype(_dtype_of(ends), copy=False))


def complement_intervals(starts, ends, bounds=(0, INT64_MAX)):
    """(starts, ends) of the gaps between the merged intervals inside `bounds`."""
    s, e = _arrays(starts, ends)
    # the reference builds its result as np.r_[bounds[0], merged ends] / np.r_[merged starts, bounds[1]]
    # (arrops.py:496-497): dtypes -- and the OverflowError for a bound that does not fit a narrow input dtype --
    # are those of that concatenation
    dt_starts = np.r_[bounds[0], np.empty(0, dtype=_dtype_of(ends))].dtype
    dt_ends = np.r_[np.empty(0, dtype=_dtype_of(starts)), bounds[1]].dtype
    if len(s) == 0:
        # an empty set has the whole bounds as its complement; the reference (arrops.py:482-503 on the numpy of
        # this image) answers the same, ([bounds[0]], [bounds[1]])
        return np.array([bounds[0]], dtype=dt_starts), np.array([bounds[1]], dtype=dt_ends)
    _, gs, ge = _engine.complement(None, s, e, np.array([bounds[0]], np.int64), np.array([bounds[1]], np.int64))
    return _host(gs).astype(dt_starts, copy=False), _host(ge).astype(dt_ends, copy=False)


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
    """(ids1, ids2) of the k closest set-2 intervals of every set-1 interval;
    with `starts2=ends2=None` the closest other intervals of the same set."""
    if tie_arr is not None:
        # np.lexsort([ids2, tie_arr, dists, ids1]) of the reference (arrops.py:738-740) mixes keys of
        # different lengths: it raises this error whenever tie_arr is given
        raise ValueError("all keys need to be the same shape")
    self_closest = starts2 is None and ends2 is None
    if self_closest:
        if ignore_overlaps:  # reference quirk: starts2 stays None and indexing fails
            raise IndexError("closest_intervals(starts1, ends1, None, None, ignore_overlaps=True) is not defined")
        s1, e1 = _arrays(starts1, ends1)
        s2 = e2 = None
    else:
        s1, e1, s2, e2 = _arrays(starts1, ends1, starts2, ends2)
    ev1, ev2, n_matched = _engine.closest_intidxs(
        None, s1, e1, None, s2, e2, 1, k=k, ignore_overlaps=ignore_overlaps, ignore_upstream=ignore_upstream,
        ignore_downstream=ignore_downstream, along=None if along is None else np.asarray(along, dtype=bool),
        self_closest=self_closest,
    )
    return _host(ev1)[:n_matched], _host(ev2)[:n_matched]
