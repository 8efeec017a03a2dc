"""Frame-level API: the host-side mirror of `bioframe.ops` for the six
north-star operations (reference src/bioframe/ops.py, v0.8.0).

Same names, arguments, defaults, errors and output frames as the reference;
the per-chromosome Python loops around `bioframe.core.arrops` are replaced by
ONE call per operation into the CUDA library (all groups at once, keyed by a
group rank), through `bioframe_b200.engine` -> C ABI.  What stays in Python is
what the reference also does in pandas: validation, chromosome encoding and the
assembly of the output frame.

  overlap / _overlap_intidxs   ops.py:361-556 / 242-338
  closest / _closest_intidxs   ops.py:1043-1240 / 919-1040
  cluster, merge               ops.py:559-708, 711-839
  coverage                     ops.py:842-916
  complement                   ops.py:1560-1687

There is no CPU fallback: without the CUDA library or a GPU these raise.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import engine as _engine
from .core import checks
from .core.specs import _get_default_colnames, _verify_columns

__all__ = ["overlap", "closest", "cluster", "merge", "complement", "coverage"]

_INT64_MAX = np.iinfo(np.int64).max


# --------------------------------------------------------------------------
# group encoding (reference: df.groupby(...).indices / .groups)
# --------------------------------------------------------------------------
def _factorize_column(col):
    """-> (codes int64 with -1 for NA, keys) with keys in the order pandas'
    sorted groupby visits them: lexical for strings, category order (observed
    only) for categoricals."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        raw = np.asarray(col.cat.codes, dtype=np.int64)
        n_cat = len(col.cat.categories)
        seen = np.zeros(n_cat + 1, dtype=bool)
        seen[raw + 1] = True
        present = np.flatnonzero(seen[1:])
        remap = np.full(n_cat + 1, -1, dtype=np.int64)
        remap[present + 1] = np.arange(len(present))
        return remap[raw + 1], [col.cat.categories[i] for i in present]
    codes, uniques = pd.factorize(col, sort=True)
    return np.asarray(codes, dtype=np.int64), list(uniques)


def _group_keys(df, by):
    """Group code per row (-1 where any key column is NA) and the list of group
    keys in sorted-groupby order.  Keys are scalars for one column and tuples
    for several, like the keys of `df.groupby(by).indices`."""
    if len(by) == 1:
        return _factorize_column(df[by[0]])
    gb = df.groupby(by, observed=True, sort=True, dropna=True)
    codes = np.asarray(gb.ngroup(), dtype=np.int64)  # -1 for rows with an NA key
    keys = list(gb.size().index)
    return codes, keys


def _int64_coords(df, col):
    """Coordinate column as int64 numpy; NA (nullable dtypes) -> 0 (such rows
    never reach a kernel as pairable: they sit in their own NA group)."""
    values = df[col]
    if values.isna().any():
        values = values.fillna(0)
    return np.asarray(values, dtype=np.int64)


def _to_host(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


# --------------------------------------------------------------------------
# overlap
# --------------------------------------------------------------------------
def _overlap_intidxs(df1, df2, how="left", cols1=None, cols2=None, on=None, _group_order=None):
    """Integer row positions of overlapping pairs, -1 where a kept row has no
    partner.  Mirrors ops._overlap_intidxs (ops.py:242-338): groups are
    (chrom, *on); the group blocks come in the iteration order of
    `set.union(set(groups1), set(groups2))` exactly as in the reference
    (ops.py:289,301), each block = pairs, unpaired-left, unpaired-right."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    _verify_columns(df1, [ck1, sk1, ek1])
    _verify_columns(df2, [ck2, sk2, ek2])
    by1, by2 = [ck1], [ck2]
    if on is not None:
        by1 = by1 + list(on)
        by2 = by2 + list(on)

    codes1, keys1 = _group_keys(df1, by1)
    codes2, keys2 = _group_keys(df2, by2)
    # dropna=False in the reference: NA-keyed rows form one more group per frame,
    # visited where the NaN key falls in the set order; they can never pair.
    na1 = bool((codes1 < 0).any())
    na2 = bool((codes2 < 0).any())
    nan_key = float("nan")
    k1 = keys1 + ([nan_key] if na1 else [])
    k2 = keys2 + ([nan_key] if na2 else [])
    order = list(set.union(set(k1), set(k2))) if _group_order is None else list(_group_order)

    # rank of every group in the visiting order; the NA groups of the two frames
    # get two adjacent ranks so that they stay unpaired
    rank = {}
    next_rank = 0
    rank_na1 = rank_na2 = -1
    for key in order:
        if isinstance(key, float) and key != key:
            rank_na1, rank_na2 = next_rank, next_rank + 1
            next_rank += 2
        else:
            rank[key] = next_rank
            next_rank += 1
    n_groups = max(next_rank, 1)

    def ranks_of(codes, keys, rank_na):
        table = np.array([rank[k] for k in keys] + [rank_na], dtype=np.int32)
        return table[codes]  # code -1 picks the last entry

    g1 = ranks_of(codes1, keys1, rank_na1) if len(df1) else np.empty(0, np.int32)
    g2 = ranks_of(codes2, keys2, rank_na2) if len(df2) else np.empty(0, np.int32)
    ev1, ev2, _ = _engine.overlap_intidxs(
        g1,
        _int64_coords(df1, sk1),
        _int64_coords(df1, ek1),
        g2,
        _int64_coords(df2, sk2),
        _int64_coords(df2, ek2),
        n_groups,
        how=how,
    )
    return _to_host(ev1), _to_host(ev2)


_NULLABLE_INT = {
    np.dtype(t): getattr(pd, n)()
    for t, n in (
        (np.int8, "Int8Dtype"),
        (np.int16, "Int16Dtype"),
        (np.int32, "Int32Dtype"),
        (np.int64, "Int64Dtype"),
        (np.uint8, "UInt8Dtype"),
        (np.uint16, "UInt16Dtype"),
        (np.uint32, "UInt32Dtype"),
        (np.uint64, "UInt64Dtype"),
    )
}


def _to_nullable_dtype(dtype):
    try:
        return _NULLABLE_INT.get(np.dtype(dtype), dtype)
    except TypeError:
        return dtype


def _suffixed_rows(df, positions, suffix):
    part = df.iloc[positions].reset_index(drop=True)
    part.columns = [c + suffix for c in part.columns]
    return part


def overlap(
    df1,
    df2,
    how="left",
    return_input=True,
    return_index=False,
    return_overlap=False,
    suffixes=("", "_"),
    keep_order=None,
    cols1=None,
    cols2=None,
    on=None,
    ensure_int=True,
):
    """Find pairs of overlapping genomic intervals (ops.py:361-556).

    Same contract as `bioframe.overlap`: how in {'left','right','outer','inner'},
    `return_index`/`return_overlap` may be column-name stems, `keep_order`
    defaults to True for how='left' and is an error otherwise, `on` adds
    grouping columns, `ensure_int` keeps coordinates integer (nullable) when
    rows go missing."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])

    if how == "left" and keep_order is None:
        keep_order = True
    if how != "left" and keep_order:
        raise ValueError("keep_order=True only allowed for how='left'")
    if on is not None:
        if not isinstance(on, list):
            raise ValueError("on=[] must be None or list")
        if (ck1 in on) or (ck2 in on):
            raise ValueError("on=[] should not contain chromosome colnames")
        _verify_columns(df1, on)
        _verify_columns(df2, on)

    ev1, ev2 = _overlap_intidxs(df1, df2, how=how, cols1=cols1, cols2=cols2, on=on)

    stem = return_index if isinstance(return_index, str) else "index"
    icol1, icol2 = stem + suffixes[0], stem + suffixes[1]
    parts = {}
    parts["i1"] = pd.DataFrame({icol1: df1.index[ev1]}, dtype=pd.Int64Dtype())
    parts["i2"] = pd.DataFrame({icol2: df2.index[ev2]}, dtype=pd.Int64Dtype())

    if return_overlap:
        ostem = return_overlap if isinstance(return_overlap, str) else "overlap"
        parts["ov"] = pd.DataFrame(
            {
                ostem + "_" + sk1: np.maximum(df1[sk1].values[ev1], df2[sk2].values[ev2]),
                ostem + "_" + ek1: np.minimum(df1[ek1].values[ev1], df2[ek2].values[ev2]),
            }
        )
    if return_input:
        parts["in1"] = _suffixed_rows(df1, ev1, suffixes[0])
        parts["in2"] = _suffixed_rows(df2, ev2, suffixes[1])

    if how != "inner":  # blank the side that has no partner (ops.py:511-544)
        miss1, miss2 = ev1 == -1, ev2 == -1
        any1, any2 = bool(miss1.any()), bool(miss2.any())
        parts["i1"][miss1] = None
        parts["i2"][miss2] = None
        for tag, miss, any_miss, df, sk, ek, suf in (
            ("in1", miss1, any1, df1, sk1, ek1, suffixes[0]),
            ("in2", miss2, any2, df2, sk2, ek2, suffixes[1]),
        ):
            if tag not in parts or not any_miss:
                continue
            if ensure_int:
                parts[tag] = parts[tag].astype(
                    {sk + suf: _to_nullable_dtype(df[sk].dtype), ek + suf: _to_nullable_dtype(df[ek].dtype)}
                )
            parts[tag][miss] = None
        if "ov" in parts and ensure_int and (any1 or any2):
            parts["ov"] = parts["ov"].convert_dtypes()
            parts["ov"][miss1 | miss2] = None

    out = pd.concat([parts.get(k) for k in ("i1", "in1", "i2", "in2", "ov")], axis=1)
    if keep_order:
        out = out.sort_values([icol1, icol2])
    if not return_index:
        out.drop([icol1, icol2], axis=1, inplace=True)
    out.reset_index(drop=True, inplace=True)
    return out
