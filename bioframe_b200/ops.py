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

import os

import numpy as np
import pandas as pd

from . import engine as _engine
from .core import checks
from .core.specs import _get_default_colnames, _verify_columns

__all__ = ["overlap", "closest", "cluster", "merge", "complement", "coverage", "count_overlaps", "setdiff", "subtract", "assign_view", "trim"]

_INT64_MAX = np.iinfo(np.int64).max

# ---- where a frame-level call spends its time (bench.py's frame_e2e leg): stage -> seconds while a
# `profile()` context is open.  Stage boundaries synchronise the device, so the context is for measuring only.
_PROFILE = None


class profile:
    def __enter__(self):
        global _PROFILE
        _PROFILE = {}
        return _PROFILE

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None
        return False


def _tick(stage, t0):
    """Charge the time since t0 to `stage`; -> now.  A no-op (returns t0) outside a profile() context."""
    if _PROFILE is None:
        return t0
    import time

    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.synchronize()
    except ImportError:
        pass
    now = time.perf_counter()
    _PROFILE[stage] = _PROFILE.get(stage, 0.0) + (now - t0)
    return now


def _now():
    if _PROFILE is None:
        return 0.0
    import time

    return time.perf_counter()
_ORDERED_MAX_PARTNERS = 1 << 14  # longest partner list of one row the row-order overlap kernel is used for


# --------------------------------------------------------------------------
# group encoding (reference: df.groupby(...).indices / .groups)
# --------------------------------------------------------------------------
def _factorize_column(col):
    """-> (codes int64 with -1 for NA, keys) with keys in the order pandas'
    sorted groupby visits them: lexical for strings, category order (observed
    only) for categoricals."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        raw = np.asarray(col.cat.codes)
        n_cat = len(col.cat.categories)
        if raw.dtype.itemsize <= 2 and len(raw):
            # 1- or 2-byte dictionary codes (the usual chromosome column): which categories occur comes from one
            # histogram over the unsigned view (-1 = NA lands in the last bin, never a category: n_cat <= 32767),
            # and when the observed categories are the first ones -- nearly always all of them -- the group codes
            # ARE the dictionary codes
            bins = np.bincount(raw.view(np.uint8 if raw.dtype.itemsize == 1 else np.uint16), minlength=n_cat)
            present = np.flatnonzero(bins[:n_cat])
            keys = [col.cat.categories[i] for i in present]
            if len(present) == 0 or present[-1] == len(present) - 1:
                return raw.astype(np.int64), keys
            remap = np.full(n_cat + 1, -1, dtype=np.int64)
            remap[present + 1] = np.arange(len(present))
            return remap[raw.astype(np.int64) + 1], keys
        raw = raw.astype(np.int64)
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
    ng = gb.ngroup().to_numpy(dtype=np.float64, na_value=np.nan)  # NaN (or -1) for rows with an NA key
    codes = np.where(np.isnan(ng) | (ng < 0), -1, ng).astype(np.int64)
    keys = list(gb.size().index)
    return codes, keys


def _is_na(x):
    return x is None or x is pd.NA or x is pd.NaT or (isinstance(x, float) and x != x)


def _visiting_order(keys1, keys2):
    """Group keys of both frames in the order the reference's loop visits them: the iteration order of
    `set.union(set(df1_groups), set(df2_groups))` (ops.py:289, 301; 985-986).  The reference builds its two sets
    from DICTS (`groupby(...).indices`), and CPython sizes a set built from a dict once, up front, while a set built
    from a list grows step by step and re-inserts its entries at every growth: with colliding hashes the two end up
    in different iteration orders.  So the keys go through dicts here too (same keys, same insertion order, same
    hashes -> the same table layout)."""
    return list(set.union(set(dict.fromkeys(keys1)), set(dict.fromkeys(keys2))))


def _chrom_is_na(key):
    """Does a group key (scalar, or tuple with the chromosome first) have a null chromosome?"""
    return _is_na(key[0]) if isinstance(key, tuple) else _is_na(key)


def _group_keys_shared_na(df, by):
    """Group code per row (always >= 0) and the keys in sorted-groupby order with NA as a key value of its
    own, the way `df.groupby(by, dropna=False).indices` of the reference sees them (ops.py:287-288): every
    null is the np.nan singleton, so a key such as ('chr1', nan) -- a null in an `on` column -- is ONE key
    shared by both frames.  Keys whose chromosome is null belong to all-null rows (checks.is_bedframe): the
    callers keep those from pairing."""
    if len(by) == 1:
        codes, keys = _factorize_column(df[by[0]])
        if (codes < 0).any():
            codes = np.where(codes < 0, len(keys), codes)
            keys = keys + [np.nan]
        return codes, keys
    gb = df.groupby(by, observed=True, sort=True, dropna=False)
    codes = gb.ngroup().to_numpy(dtype=np.int64)
    keys = [tuple(np.nan if _is_na(x) else x for x in key) for key in gb.size().index]
    return codes, keys


# measured on the B200 box (profiles/threshold_sweep.py): the device encoding wins from 1e5 rows up (3.4 vs 3.9 ms for a
# whole 1e5 x 1e5 inner join, 37 vs 52 ms at 1e6 x 1e6)
_DEVICE_ENCODE_MIN_ROWS = 1 << 16


def _categorical_keys_on_device(df1, by1, df2, by2):
    """Group encoding of two frames keyed by ONE categorical column each (the usual chromosome column) without
    a pass over the rows on the host: the dictionary codes (int8 / int16) go to the device as they are,
    bf_group_codes_present reports the observed categories -- the keys `groupby(observed=True)` of the reference
    would find (ops.py:287-288), nulls as the np.nan key -- and the caller maps codes to group ranks there too.
    -> ((device codes, keys, lookup slots) per frame) or None when the frames are keyed otherwise / are small."""
    if len(by1) != 1 or len(by2) != 1 or not hasattr(_engine, "codes_present"):
        return None
    if max(len(df1), len(df2)) < _DEVICE_ENCODE_MIN_ROWS or min(len(df1), len(df2)) == 0:
        return None
    out = []
    for df, col in ((df1, by1[0]), (df2, by2[0])):
        dtype = df[col].dtype
        if not isinstance(dtype, pd.CategoricalDtype):
            return None
        raw = df[col].cat.codes.to_numpy()
        if raw.dtype.itemsize > 2:
            return None
        dev, present = _engine.codes_present(raw, len(dtype.categories))
        slots = np.flatnonzero(present[1:]) + 1  # observed categories, in category order like the sorted groupby
        keys = [dtype.categories[i - 1] for i in slots]
        if present[0]:  # nulls: one more key, after the others (_group_keys_shared_na)
            keys.append(np.nan)
            slots = np.append(slots, 0)
        out.append((dev, keys, slots))
    return out


def _int64_coords(df, col):
    """Coordinate column as int64 numpy; NA (nullable dtypes) -> 0 (such rows
    never reach a kernel as pairable: they sit in their own NA group)."""
    values = df[col]
    if isinstance(values.dtype, np.dtype) and values.dtype.kind in "iu":
        return np.asarray(values, dtype=np.int64)  # a plain integer column cannot hold a null
    if values.isna().any():
        values = values.fillna(0)
    return np.asarray(values, dtype=np.int64)


def _ranked_rows(df, codes, table, sk, ek, keep=None):
    """The rows of a frame that take part in a kernel call: those with a group code (>= 0), or, with `keep` (a
    flag per group key), those of a kept group.  `table` maps group codes to group ranks (int32).
    -> (row positions, or None when that is every row; group ranks; starts; ends).  The usual frame has no null
    and its codes already are the ranks: then nothing is indexed or copied per row beyond the int32 cast."""
    n = len(codes)
    if keep is not None:
        rows = None if bool(keep.all()) else np.flatnonzero(keep[codes])
    else:
        rows = None if (n == 0 or codes.min() >= 0) else np.flatnonzero(codes >= 0)
    identity = np.array_equal(table, np.arange(len(table)))
    starts, ends = _int64_coords(df, sk), _int64_coords(df, ek)
    if rows is None:
        return None, (codes.astype(np.int32) if identity else table[codes]), starts, ends
    picked = codes[rows]
    return rows, (picked.astype(np.int32) if identity else table[picked]), starts[rows], ends[rows]


def _spread(values, rows, n, fill=0):
    """values of the participating rows -> one value per frame row (`fill` elsewhere)."""
    if rows is None:
        return values
    out = np.full(n, fill, dtype=values.dtype)
    out[rows] = values
    return out


def _to_host(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def _rows_to_host(t):
    """Row-number results (int64, -1 = none): large ones cross PCIe as 4 bytes per value and are widened on
    the host (bf_download_rows); small ones take the plain copy."""
    if hasattr(t, "is_cuda") and t.is_cuda and t.numel() >= (1 << 22):
        return _engine.download_rows(t)
    return _to_host(t)


# --------------------------------------------------------------------------
# rows the reference's grouping leaves out
# --------------------------------------------------------------------------
_PANDAS_DROPS_NULL_CATEGORY = None


def _pandas_drops_null_category():
    """Does `groupby([one categorical column], observed=True, dropna=False).indices` of this pandas leave the
    null rows out?  (It does in pandas 1.x-3.x: the single-categorical fast path of `.indices` knows no NaN
    group, although `dropna=False` -- with two key columns, or an object column, the NaN key is there.)  The
    reference builds its groups with exactly that call (ops.py:287-288), so where pandas drops them, all-null
    rows of a frame with a categorical chromosome column appear nowhere in its `_overlap_intidxs` result -- not
    even as unpaired rows of a left / outer join.  Probed once instead of assumed."""
    global _PANDAS_DROPS_NULL_CATEGORY
    if _PANDAS_DROPS_NULL_CATEGORY is None:
        probe = pd.DataFrame({"k": pd.Categorical(["a", None])})
        _PANDAS_DROPS_NULL_CATEGORY = len(probe.groupby(["k"], observed=True, dropna=False).indices) == 1
    return _PANDAS_DROPS_NULL_CATEGORY


def _ungrouped_rows(df, ck, on):
    """Mask of the rows `_overlap_intidxs` of the reference never sees (see above), or None."""
    if on or not isinstance(df[ck].dtype, pd.CategoricalDtype):
        return None
    missing = np.asarray(df[ck].cat.codes) < 0
    if not missing.any() or not _pandas_drops_null_category():
        return None
    return missing


def _frame_positions(ev, rows):
    """Row numbers of a filtered frame -> row numbers of the frame it was taken from (-1 stays)."""
    if rows is None:
        return ev
    return np.where(ev >= 0, rows[np.maximum(ev, 0)], -1) if len(rows) else ev


# --------------------------------------------------------------------------
# overlap
# --------------------------------------------------------------------------
def _encode_frames(df1, df2, cols1=None, cols2=None, on=None, _group_order=None, _device=True):
    """The two frames as the arrays the engine takes: group rank per row (int32; position of the row's
    (chrom, *on) group in the reference's visiting order, ops.py:287-301) and int64 coordinates.
    -> (g1, s1, e1, g2, s2, e2, n_groups, order); `order` is the list of group keys in visiting order (the order
    of the union set, which follows the process's hash seed for strings: a multi-process caller fixes it by
    passing one rank's `order` to the others as _group_order).  _device=False keeps the ranks on the host."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    _verify_columns(df1, [ck1, sk1, ek1])
    _verify_columns(df2, [ck2, sk2, ek2])
    by1, by2 = [ck1], [ck2]
    if on is not None:
        by1 = by1 + list(on)
        by2 = by2 + list(on)

    fast = _categorical_keys_on_device(df1, by1, df2, by2) if _device else None
    if fast is not None:
        (dev1, keys1, slots1), (dev2, keys2, slots2) = fast
    else:
        codes1, keys1 = _group_keys_shared_na(df1, by1)
        codes2, keys2 = _group_keys_shared_na(df2, by2)
    # dropna=False in the reference (ops.py:287-288): a null in an `on` column is a key value like any other,
    # shared by both frames; the groups are visited in the order of the union set (ops.py:289, 301)
    order = _visiting_order(keys1, keys2) if _group_order is None else [
        np.nan if _is_na(k) else k for k in _group_order]

    # rank of every group in the visiting order.  A key with a null CHROMOSOME holds all-null rows
    # (checks.is_bedframe): their coordinates are NaN in the reference and never overlap anything, so the two
    # frames' rows of such a key get two adjacent ranks -- same block position, never paired.
    rank1, rank2 = {}, {}
    next_rank = 0
    for key in order:
        if _chrom_is_na(key):
            rank1[key], rank2[key] = next_rank, next_rank + 1
            next_rank += 2
        else:
            rank1[key] = rank2[key] = next_rank
            next_rank += 1
    n_groups = max(next_rank, 1)

    def ranks_of(codes, keys, rank):
        table = np.array([rank[k] for k in keys], dtype=np.int32)
        return table[codes]

    if fast is not None:  # dictionary codes -> group ranks on the device (bf_group_codes_apply)
        def lut_of(keys, slots, rank, n_slots):
            lut = np.full(n_slots, -1, dtype=np.int32)
            lut[slots] = [rank[k] for k in keys]
            return lut

        g1 = _engine.codes_apply(dev1, lut_of(keys1, slots1, rank1, len(df1[by1[0]].cat.categories) + 1))
        g2 = _engine.codes_apply(dev2, lut_of(keys2, slots2, rank2, len(df2[by2[0]].cat.categories) + 1))
    else:
        g1 = ranks_of(codes1, keys1, rank1) if len(df1) else np.empty(0, np.int32)
        g2 = ranks_of(codes2, keys2, rank2) if len(df2) else np.empty(0, np.int32)
    s1, e1, s2, e2 = _int64_coords(df1, sk1), _int64_coords(df1, ek1), _int64_coords(df2, sk2), _int64_coords(df2, ek2)
    return g1, s1, e1, g2, s2, e2, n_groups, order


def _overlap_intidxs(df1, df2, how="left", cols1=None, cols2=None, on=None, _group_order=None, _sort_pairs=False,
                     _coords=False, _device_rows=None):
    """Integer row positions of overlapping pairs, -1 where a kept row has no
    partner.  Mirrors ops._overlap_intidxs (ops.py:242-338): groups are
    (chrom, *on); the group blocks come in the iteration order of
    `set.union(set(groups1), set(groups2))` exactly as in the reference
    (ops.py:289,301), each block = pairs, unpaired-left, unpaired-right."""
    t0 = _now()
    gone1 = _ungrouped_rows(df1, (cols1 or _get_default_colnames())[0], on)
    gone2 = _ungrouped_rows(df2, (cols2 or _get_default_colnames())[0], on)
    if gone1 is not None or gone2 is not None:
        # all-null rows under a categorical chromosome column: the reference's grouping has no group for them
        # (_pandas_drops_null_category), so they are not even unpaired rows.  The join runs on the other rows.
        rows1 = None if gone1 is None else np.flatnonzero(~gone1)
        rows2 = None if gone2 is None else np.flatnonzero(~gone2)
        res = _overlap_intidxs(df1 if rows1 is None else df1.iloc[rows1], df2 if rows2 is None else df2.iloc[rows2],
                               how=how, cols1=cols1, cols2=cols2, on=on, _group_order=_group_order,
                               _sort_pairs=_sort_pairs, _coords=_coords)
        return (_frame_positions(res[0], rows1), _frame_positions(res[1], rows2), *res[2:])
    g1, s1, e1, g2, s2, e2, n_groups, _order = _encode_frames(df1, df2, cols1, cols2, on, _group_order)
    t0 = _tick("encode_groups", t0)
    if _coords or _PROFILE is not None:
        # the coordinates go to the device once: the overlap call and the coordinate gather share them
        s1, e1, s2, e2 = (_engine.to_device(x) for x in (s1, e1, s2, e2))
        g1, g2 = _engine.to_device(g1, "int32"), _engine.to_device(g2, "int32")
    t0 = _tick("h2d", t0)
    ev1 = None
    if _sort_pairs and how == "left":
        # keep_order: the pairs are produced in (row of df1, row of df2) order directly, queries visited in row
        # order (bf_overlap_ordered_*).  A row with a huge partner list is sorted by one thread there: such
        # inputs take the general path below (all pairs, then a radix sort of the pair list).
        try:
            ev1, ev2, _biggest = _engine.overlap_ordered(g1, s1, e1, g2, s2, e2, n_groups,
                                                         max_partners=_ORDERED_MAX_PARTNERS)
        except OverflowError:  # the key layout of all groups together does not fit: the general path batches them
            ev1 = None
    if ev1 is None:
        # any other string joins like 'inner' in the reference (its `how in {...}` tests all fail, ops.py:229-234,
        # 327-331; its own test suite passes how="innter")
        join = how if how in ("left", "right", "outer") else "inner"
        ev1, ev2, _ = _engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, n_groups, how=join)
        if _sort_pairs:  # keep_order on the device: rows by (row of df1, row of df2), -1 last
            ev1, ev2 = _engine.sort_pairs(ev1, ev2, max(len(df1) - 1, 0), max(len(df2) - 1, 0))
    if _device_rows is not None:  # the caller goes on working with the pair list where it is
        _device_rows.extend([ev1, ev2])
    if _coords:  # clipped coordinates of every pair (return_overlap, ops.py:480-497) while the list is on the device
        lo, hi, _ = _engine.pair_coords(ev1, ev2, s1, e1, s2, e2)
        t0 = _tick("kernels", t0)
        out = _rows_to_host(ev1), _rows_to_host(ev2), (_to_host(lo), _to_host(hi))
        _tick("d2h", t0)
        return out
    t0 = _tick("kernels", t0)
    out = _rows_to_host(ev1), _rows_to_host(ev2)
    _tick("d2h", t0)
    return out


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


def _is_plain_int(col):
    return isinstance(col.dtype, np.dtype) and col.dtype.kind in "iu" and col.dtype.itemsize <= 8 and col.dtype != np.uint64


def _to_nullable_dtype(dtype):
    try:
        return _NULLABLE_INT.get(np.dtype(dtype), dtype)
    except TypeError:
        return dtype


def _int64_array(values):
    """A nullable Int64 column without a null over `values` (int64 numpy, not copied)."""
    return pd.arrays.IntegerArray(np.ascontiguousarray(values, dtype=np.int64), np.zeros(len(values), dtype=bool))


def _suffixed_rows(df, positions, suffix):
    part = df.iloc[positions].reset_index(drop=True)
    if suffix != "":
        part.columns = [c + suffix for c in part.columns]
    return part


# Results of this many rows and more are assembled with device-side column gathers (bf_gather_rows): taking
# 1e8+ rows of a column with numpy costs seconds per column; the device does it at PCIe speed.  The source columns
# have to cross PCIe first, so small results stay with pandas: measured (profiles/threshold_sweep.py) a loss at
# 5.8e4 result rows of 3e5-row frames (4.6 vs 2.5 ms), a gain from 5.8e5 result rows up (27 vs 33 ms; 1e6 x 1e6:
# 15 vs 31 ms; 1e7 x 1e6: 118 vs 459 ms).
_DEVICE_GATHER_MIN_ROWS = 1 << 19
_SAME_WIDTH_INT = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}


def _device_take(values, device_positions):
    """values[positions] for a fixed-width numpy array (1, 2, 4 or 8 bytes per item, any dtype of that width)
    through bf_gather_rows; position -1 yields 0.  -> numpy array of the same dtype."""
    values = np.ascontiguousarray(values)
    as_int = _SAME_WIDTH_INT[values.dtype.itemsize]
    taken = _engine.gather_rows(values.view(as_int), device_positions)
    small = values.dtype == np.int64 and len(values) and values.min() >= -1 and values.max() < 2**31 - 1
    host = _rows_to_host(taken) if small else _to_host(taken)
    return host.view(values.dtype)


def _fixed_width(dtype):
    return isinstance(dtype, np.dtype) and dtype.kind in "iufbmM" and dtype.itemsize in _SAME_WIDTH_INT


def _taken_column(col, positions, device_positions):
    """One column of `df.iloc[positions]` (ops.py:499-508) as a pandas array, built from device gathers where
    the column has a fixed-width representation -- plain numeric / bool / datetime values, the integer codes of
    a categorical (chromosome names!), the value + mask pair of a nullable column (the Arrow layout) -- and
    with pandas' take otherwise (object / string columns).  Position -1 picks an arbitrary valid value: the
    caller blanks those rows, as the reference blanks what it took from position -1."""
    dtype = col.dtype
    if _fixed_width(dtype):
        return _device_take(col.to_numpy(), device_positions)
    if isinstance(dtype, pd.CategoricalDtype) and len(dtype.categories):
        codes = _device_take(np.asarray(col.cat.codes), device_positions)
        return pd.Categorical.from_codes(codes, dtype=dtype)
    arr = col.array
    if isinstance(arr, (pd.arrays.IntegerArray, pd.arrays.FloatingArray, pd.arrays.BooleanArray)) and _fixed_width(arr._data.dtype):
        values = _device_take(arr._data, device_positions)
        mask = _device_take(arr._mask, device_positions)
        return type(arr)(values, mask)
    if isinstance(arr, pd.arrays.ArrowStringArray):
        taken = _arrow_string_take(arr, positions, device_positions)
        if taken is not None:
            return taken
    return col.iloc[positions].array


_ARROW_TAKE_THREADS = max(1, min(8, os.cpu_count() or 1))


# A take of this many values and more out of a string column with few distinct values (names of regions, genes,
# samples ...) goes through the column's dictionary when every row is taken at least once on average: see
# _arrow_string_take.  _DICT_TAKE_ALWAYS: tests.
_DICT_TAKE_MIN_ROWS = 1 << 20
_DICT_TAKE_ALWAYS = False


def _threaded(fn, parts):
    if parts == 1:
        return [fn(0)]
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(parts) as pool:
        return list(pool.map(fn, range(parts)))


def _string_dictionary(values, parts):
    """(int32 code per row, -1 for nulls; dictionary) of an Arrow string array with few distinct values, encoded
    slice by slice on `parts` threads (pyarrow releases the GIL) and unified; None when the column has too many
    distinct values for a dictionary to pay (decided on its first 65 536 rows, and again on the whole)."""
    import pyarrow as pa
    import pyarrow.compute as pc

    n = len(values)
    if len(pc.dictionary_encode(values.slice(0, 1 << 16)).dictionary) > (1 << 12):
        return None
    bounds = np.linspace(0, n, parts + 1).astype(np.int64)

    def encode(k):
        piece = pc.dictionary_encode(values.slice(int(bounds[k]), int(bounds[k + 1] - bounds[k])))
        return piece.dictionary, pc.fill_null(piece.indices, -1).to_numpy(zero_copy_only=False)

    pieces = _threaded(encode, parts)
    dictionary = pc.unique(pa.concat_arrays([d for d, _ in pieces]))
    if len(dictionary) == 0 or len(dictionary) > max(n >> 4, 1 << 12):
        return None
    codes = np.empty(n, dtype=np.int32)

    def unify(k):
        local_dictionary, local_codes = pieces[k]
        table = np.append(pc.index_in(local_dictionary, value_set=dictionary).to_numpy(zero_copy_only=False), -1)
        codes[bounds[k]:bounds[k + 1]] = table.astype(np.int32)[local_codes]  # local code -1 (null) -> last entry, -1

    _threaded(unify, parts)
    return codes, dictionary


def _arrow_string_take(arr, positions, device_positions=None):
    """`arr.take(positions)` for an Arrow-backed string column, split over a few host threads (pyarrow's take
    releases the GIL): variable-width values cannot be gathered on the device without shipping the bytes both
    ways, and pandas' single-threaded `iloc` of a string column was the largest share of a 1e7 x 1e6 inner
    join's assembly.  Position -1 yields a missing (or arbitrary) value: the caller blanks those rows anyway.
    None when this pandas / pyarrow does not expose what is needed -- the caller falls back to `iloc`.

    A take at random positions of a large column is bound by cache misses (290 ms on one thread for 6.5e6 of 1e7
    values, 50 ms for the same positions sorted).  When a column with few distinct values is taken more than once
    over (joins with fan-out), it is dictionary-encoded instead (a sequential pass, sliced over the threads and
    unified: ~200 ms per 1e7 rows on one thread), the 4-byte codes are gathered on the device like any fixed-width
    column, and the result is decoded from the small, cache-resident dictionary (~10 ns a value instead of 25-45)."""
    try:
        import pyarrow as pa
        import pyarrow.compute as pc

        values = arr._pa_array
        if isinstance(values, pa.ChunkedArray):
            values = values.combine_chunks() if values.num_chunks != 1 else values.chunk(0)
    except (ImportError, AttributeError):
        return None
    positions = np.asarray(positions, dtype=np.int64)
    parts = max(1, min(_ARROW_TAKE_THREADS, len(positions) >> 18))
    bounds = np.linspace(0, len(positions), parts + 1).astype(np.int64)
    # worth it when the sequential pass over the column (~20 ns a row: encode, codes, unify) costs less than what the
    # decode saves per taken row (25 ns from a 1e6-row column, 45 ns from a 1e7-row one, against ~10 ns from the
    # dictionary): a column taken at least once over (twice for smaller columns), i.e. joins with fan-out
    n_taken, n_rows = len(positions), len(values)
    through_dictionary = _DICT_TAKE_ALWAYS or (
        n_rows >= (1 << 16) and n_taken >= _DICT_TAKE_MIN_ROWS and n_taken >= (n_rows if n_rows >= (1 << 22) else 2 * n_rows))
    if device_positions is not None and through_dictionary:
        encoded = _string_dictionary(values, max(1, min(_ARROW_TAKE_THREADS, len(values) >> 18)))
        if encoded is not None:
            codes, dictionary = encoded
            taken = _device_take(codes, device_positions)  # position -1 -> code 0, a valid entry (blanked later)

            def decode(k):
                rows = taken[bounds[k]:bounds[k + 1]]
                missing = rows < 0
                rows = pa.array(rows, mask=missing) if missing.any() else pa.array(rows)
                return pc.take(dictionary, rows, boundscheck=False)

            return type(arr)(pa.chunked_array(_threaded(decode, parts), type=values.type), dtype=arr.dtype)

    def take(k):
        rows = positions[bounds[k]:bounds[k + 1]]
        missing = rows < 0
        rows = pa.array(rows, mask=missing) if missing.any() else pa.array(rows)
        return pc.take(values, rows, boundscheck=False)

    return type(arr)(pa.chunked_array(_threaded(take, parts), type=values.type), dtype=arr.dtype)


def _taken_rows(df, positions, device_positions, suffix):
    """`df.iloc[positions]` with suffixed column names (ops.py:499-508).  For large results the pair list still
    lives on the device and every column with a fixed-width representation is gathered there
    (`_taken_column`); only object / string columns go through pandas."""
    if device_positions is None or len(positions) < _DEVICE_GATHER_MIN_ROWS or not df.columns.is_unique or len(df) == 0:
        return _suffixed_rows(df, positions, suffix)  # (an empty frame has nothing to gather from: pandas' IndexError)
    data = {c + suffix: _taken_column(df[c], positions, device_positions) for c in df.columns}
    return pd.DataFrame(data, columns=[c + suffix for c in df.columns], copy=False)


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
    t0 = _now()
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])
    _tick("validate", t0)

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

    # keep_order=True sorts the result by (index1, index2) labels (ops.py:549-550).  For strictly
    # increasing indexes that is the order of the row positions, which the device produces directly
    # (bf_sort_pairs) -- pandas' sort_values over the assembled frame dominates the reference's run time.
    device_order = bool(keep_order) and all(
        ix.is_monotonic_increasing and ix.is_unique for ix in (df1.index, df2.index)
    )
    # return_overlap gathers four coordinates per pair (ops.py:484-492): done on the device for plain integer
    # columns; nullable columns keep the host path (their NA propagation decides the reference's dtypes)
    device_coords = (
        bool(return_overlap)
        and (how == "inner" or ensure_int)
        and all(_is_plain_int(df[c]) for df, c in ((df1, sk1), (df1, ek1), (df2, sk2), (df2, ek2)))
    )
    dev_rows = []
    res = _overlap_intidxs(df1, df2, how=how, cols1=cols1, cols2=cols2, on=on, _sort_pairs=device_order,
                           _coords=device_coords, _device_rows=dev_rows)
    ev1, ev2 = res[0], res[1]
    dev1, dev2 = dev_rows if dev_rows else (None, None)
    t0 = _now()

    stem = return_index if isinstance(return_index, str) else "index"
    icol1, icol2 = stem + suffixes[0], stem + suffixes[1]
    parts = {}
    # the index columns are only built when they are returned or needed to order the rows (the reference always
    # builds them, ops.py:464-478, and drops them at the end when return_index is False, ops.py:552-553)
    need_index = bool(return_index) or (bool(keep_order) and not device_order)
    if not all(pd.api.types.is_integer_dtype(ix.dtype) for ix in (df1.index, df2.index)):
        need_index = True  # labels that are not integers: cast (or refused) exactly as the reference does below
    if (len(df1) == 0 or len(df2) == 0) and len(ev1):
        need_index = True  # "no partner" (-1) looked up in an empty index: the reference's IndexError (ops.py:464-478)
    if not (return_input or return_overlap):
        need_index = True  # nothing but the (possibly dropped) label columns: the frame they leave behind, as built there
    if need_index:
        parts["i1"] = pd.DataFrame({icol1: df1.index[ev1]}, dtype=pd.Int64Dtype())
        parts["i2"] = pd.DataFrame({icol2: df2.index[ev2]}, dtype=pd.Int64Dtype())

    if return_overlap:
        ostem = return_overlap if isinstance(return_overlap, str) else "overlap"
        if device_coords:
            ov_lo, ov_hi = res[2]
            dt = np.result_type(df1[sk1].dtype, df2[sk2].dtype), np.result_type(df1[ek1].dtype, df2[ek2].dtype)
            ov_lo, ov_hi = ov_lo.astype(dt[0], copy=False), ov_hi.astype(dt[1], copy=False)
        else:
            ov_lo = np.maximum(df1[sk1].values[ev1], df2[sk2].values[ev2])
            ov_hi = np.minimum(df1[ek1].values[ev1], df2[ek2].values[ev2])
        parts["ov"] = pd.DataFrame({ostem + "_" + sk1: ov_lo, ostem + "_" + ek1: ov_hi})
    if return_input:
        parts["in1"] = _taken_rows(df1, ev1, dev1, suffixes[0])
        parts["in2"] = _taken_rows(df2, ev2, dev2, suffixes[1])
    del dev_rows, dev1, dev2

    if how != "inner":  # blank the side that has no partner (ops.py:511-544)
        miss1, miss2 = ev1 == -1, ev2 == -1
        any1, any2 = bool(miss1.any()), bool(miss2.any())
        if need_index:
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

    kept = [parts[k] for k in ("i1", "in1", "i2", "in2", "ov") if k in parts]
    out = pd.concat(kept, axis=1)
    if keep_order and not device_order:
        out = out.sort_values([icol1, icol2])
    if need_index and not return_index:
        out.drop([icol1, icol2], axis=1, inplace=True)
    out.reset_index(drop=True, inplace=True)
    _tick("assemble_frame", t0)
    return out


# --------------------------------------------------------------------------
# closest
# --------------------------------------------------------------------------
def _closest_intidxs(
    df1,
    df2=None,
    k=1,
    ignore_overlaps=False,
    ignore_upstream=False,
    ignore_downstream=False,
    direction_col=None,
    tie_breaking_col=None,
    cols1=None,
    cols2=None,
    _group_order=None,
    _coords=None,
    _device_rows=None,
):
    """Integer row positions of the k closest intervals, -1 where a query has
    none.  Mirrors ops._closest_intidxs (ops.py:919-1040): chromosome blocks in
    the iteration order of `set.union(set(groups1), set(groups2))`
    (ops.py:985,997); rows with a null chromosome take no part (groupby drops
    them).  `tie_breaking_col` is not supported: the reference's own path for it
    fails whenever the candidate count differs from the target count
    (arrops.py:740)."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    if tie_breaking_col is not None:
        # the reference hands the column to np.lexsort next to keys of another length (arrops.py:738-740)
        # and fails with exactly this error whenever the argument is used
        raise ValueError("all keys need to be the same shape")
    self_closest = (df2 is None) or (df2 is df1)
    if self_closest:
        if len(df1) == 1:
            raise ValueError("df1 must have more than one interval to find closest non-identical interval")
        df2 = df1

    codes1, keys1 = _group_keys(df1, [ck1])
    codes2, keys2 = (codes1, keys1) if self_closest else _group_keys(df2, [ck2])
    order = _visiting_order(keys1, keys2) if _group_order is None else list(_group_order)
    rank = {key: r for r, key in enumerate(order)}
    n_groups = max(len(order), 1)

    def coded(df, codes, keys, sk, ek):
        """rows with a valid chromosome: (positions or None for every row, group ranks, starts, ends)"""
        return _ranked_rows(df, codes, np.array([rank[key] for key in keys], dtype=np.int32), sk, ek)

    rows1, g1, s1, e1 = coded(df1, codes1, keys1, sk1, ek1)
    if _coords:
        s1, e1 = _engine.to_device(s1), _engine.to_device(e1)
    along = None
    if direction_col is not None:
        along = df1[direction_col].values != "-"
        if rows1 is not None:
            along = along[rows1]
    if self_closest:
        rows2, s2, e2 = rows1, s1, e1
        ev1, ev2, _ = _engine.closest_intidxs(
            g1, s1, e1, None, None, None, n_groups, k=k, ignore_overlaps=ignore_overlaps,
            ignore_upstream=ignore_upstream, ignore_downstream=ignore_downstream, along=along, self_closest=True,
        )
    else:
        rows2, g2, s2, e2 = coded(df2, codes2, keys2, sk2, ek2)
        if _coords:
            s2, e2 = _engine.to_device(s2), _engine.to_device(e2)
        ev1, ev2, _ = _engine.closest_intidxs(
            g1, s1, e1, g2, s2, e2, n_groups, k=k, ignore_overlaps=ignore_overlaps,
            ignore_upstream=ignore_upstream, ignore_downstream=ignore_downstream, along=along,
        )
    extra = None
    if _coords:  # (overlap_start, overlap_end, distance) of every reported pair, gathered on the device
        extra = tuple(
            None if x is None else _to_host(x)
            for x in _engine.pair_coords(ev1, ev2, s1, e1, s2, e2, want_overlap="overlap" in _coords,
                                         want_distance="distance" in _coords)
        )
    if _device_rows is not None and rows1 is None and rows2 is None:
        _device_rows.extend([ev1, ev2])  # frame positions as they are: the caller gathers its columns with them
    ev1, ev2 = _rows_to_host(ev1), _rows_to_host(ev2)
    # kernel rows index the NA-free subsets; map back to frame positions
    if rows1 is not None:
        ev1 = rows1[ev1]
    ev2 = _frame_positions(ev2, rows2)  # (-1 stays; also when no row of df2 took part)
    return (ev1, ev2, extra) if _coords else (ev1, ev2)


def closest(
    df1,
    df2=None,
    k=1,
    ignore_overlaps=False,
    ignore_upstream=False,
    ignore_downstream=False,
    direction_col=None,
    tie_breaking_col=None,
    return_input=True,
    return_index=False,
    return_distance=True,
    return_overlap=False,
    suffixes=("", "_"),
    cols1=None,
    cols2=None,
):
    """For every interval of df1 the k closest intervals of df2 (ops.py:1043-1240).
    Same arguments, errors and output columns as `bioframe.closest`."""
    if k < 1:
        raise ValueError("k>=1 required")
    if df2 is df1:
        raise ValueError("pass df2=None to find closest non-identical intervals within the same set.")
    if df2 is None:
        df2 = df1
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])

    want = tuple(t for t, on_ in (("overlap", return_overlap), ("distance", return_distance)) if on_)
    dev_rows = []  # the pair list where the kernel left it: large results gather their columns there (_taken_rows)
    res = _closest_intidxs(
        df1, df2, k=k, ignore_overlaps=ignore_overlaps, ignore_upstream=ignore_upstream,
        ignore_downstream=ignore_downstream, direction_col=direction_col, tie_breaking_col=tie_breaking_col,
        cols1=cols1, cols2=cols2, _coords=want or None, _device_rows=dev_rows,
    )
    ev1, ev2 = res[0], res[1]
    dev1, dev2 = dev_rows if dev_rows else (None, None)
    del dev_rows
    lone = ev2 == -1

    pieces = []
    if return_index:
        stem = return_index if isinstance(return_index, str) else "index"
        idx2 = pd.DataFrame({stem + suffixes[1]: df2.index[ev2]})
        idx2[lone] = pd.NA
        pieces.append(("i1", pd.DataFrame({stem + suffixes[0]: df1.index[ev1]})))
        pieces.append(("i2", idx2))
    if return_input:
        in2 = _taken_rows(df2, ev2, dev2, suffixes[1]).astype(
            {sk2 + suffixes[1]: pd.Int64Dtype(), ek2 + suffixes[1]: pd.Int64Dtype()}
        )
        in2[lone] = pd.NA
        pieces.append(("in1", _taken_rows(df1, ev1, dev1, suffixes[0])))
        pieces.append(("in2", in2))
    del dev1, dev2
    if return_overlap:
        lo, hi = res[2][0], res[2][1]
        have = lo < hi
        ov = pd.DataFrame(
            {"have_overlap": have, "overlap_start": np.where(have, lo, None), "overlap_end": np.where(have, hi, None)}
        ).astype({"have_overlap": pd.BooleanDtype(), "overlap_start": pd.Int64Dtype(), "overlap_end": pd.Int64Dtype()})
        ov[lone] = pd.NA
        pieces.append(("ov", ov))
    if return_distance:
        dist = pd.DataFrame({"distance": res[2][2]}, dtype=pd.Int64Dtype())
        dist[lone] = pd.NA
        pieces.append(("d", dist))
    by_tag = dict(pieces)
    return pd.concat([by_tag.get(t) for t in ("i1", "in1", "i2", "in2", "ov", "d")], axis=1)


# --------------------------------------------------------------------------
# cluster / merge
# --------------------------------------------------------------------------
def _check_min_dist_and_on(df, min_dist, ck, on):
    if min_dist is not None and min_dist < 0:
        raise ValueError("min_dist>=0 currently required")
    group_list = [ck]
    if on is not None:
        if not isinstance(on, list):
            raise ValueError("on=[] must be None or list")
        if ck in on:
            raise ValueError("on=[] should not contain chromosome colnames")
        _verify_columns(df, on)
        group_list = group_list + on
    return group_list


def _cannot_hold_null(dtype):
    """Plain integer / bool numpy columns have no null: no pass over the rows is needed to know it."""
    return isinstance(dtype, np.dtype) and dtype.kind in "iub"


def _null_rows(df, cols):
    """Boolean mask of the rows with a null in any of `cols` (pd.isnull(df[cols]).any(axis=1), ops.py:676), or
    None when there is none; columns that cannot hold a null are not scanned."""
    mask = None
    for c in cols:
        col = df[c]
        if _cannot_hold_null(col.dtype):
            continue
        if isinstance(col.dtype, pd.CategoricalDtype):
            m = np.asarray(col.cat.codes) < 0
        else:
            m = np.asarray(pd.isnull(col))
        if m.any():
            mask = m if mask is None else (mask | m)
    return mask


def _merge_groups(df, group_list, sk, ek, min_dist, want_ids, want_row_spans=False):
    """One kernel call for all (chrom, *on) groups in sorted key order
    (`df.groupby(group_list, observed=True).groups`, ops.py:631/772).
    -> (valid row positions or None for "every row", engine result, group keys, mask of null rows or None)."""
    codes, keys = _group_keys(df, group_list)
    null_rows = _null_rows(df, [sk, ek, *group_list])
    md = None if min_dist is None else int(np.floor(min_dist))  # integer coordinates: s > r + d <=> s > r + floor(d)
    grp, starts, ends = codes.astype(np.int32), _int64_coords(df, sk), _int64_coords(df, ek)
    valid = None
    if null_rows is not None:  # rows with a null key or coordinate take no part (ops.py:631, 641-644)
        valid = np.flatnonzero(~null_rows)
        grp, starts, ends = grp[valid], starts[valid], ends[valid]
    res = _engine.cluster(grp, starts, ends, max(len(keys), 1), min_dist=md, want_ids=want_ids, want_table=True,
                          want_row_spans=want_row_spans)
    return valid, res, keys, null_rows


def _cluster_table(df, res, keys, group_list, sk, ek):
    """Frame of merged intervals: group columns, start, end, n_intervals (ops.py:663-674)."""
    cgrp = _to_host(res["cgrp"]).astype(np.int64)
    out = {}
    for pos, col in enumerate(group_list):
        # the key of every cluster's group, vectorised: the few distinct keys are looked up by group code (a Python
        # loop over a million clusters cost more than the clustering itself -- "storing chromosome names causes a
        # 2x slowdown", ops.py:660)
        per_group = [k if len(group_list) == 1 else k[pos] for k in keys]
        dtype = df[col].dtype
        if isinstance(dtype, pd.CategoricalDtype) and len(per_group):
            codes = dtype.categories.get_indexer(per_group)
            out[col] = pd.Series(pd.Categorical.from_codes(codes[cgrp], dtype=dtype))
        else:
            table = np.empty(len(per_group), dtype=object)
            table[:] = per_group
            out[col] = pd.Series(table[cgrp] if len(per_group) else [], dtype=dtype)
    out[sk] = _to_host(res["cstart"])
    out[ek] = _to_host(res["cend"])
    out["n_intervals"] = _to_host(res["ccount"])
    return pd.DataFrame(out)


def cluster(
    df,
    min_dist=0,
    cols=None,
    on=None,
    return_input=True,
    return_cluster_ids=True,
    return_cluster_intervals=True,
):
    """Cluster overlapping intervals (ops.py:559-708).  Cluster ids follow the
    groups in sorted key order; rows with a null chromosome/coordinate get one
    cluster each after all others (ops.py:676-685)."""
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    if min_dist is not None and min_dist < 0:
        raise ValueError("min_dist>=0 currently required")
    _verify_columns(df, [ck, sk, ek])
    df = df.reset_index(drop=True)
    group_list = _check_min_dist_and_on(df, min_dist, ck, on)
    if len(df) == 0:
        raise ValueError("No objects to concatenate")  # pd.concat([]) over the reference's cluster table (ops.py:684)

    # every row's cluster span comes straight from the device (row_start / row_end of bf_cluster_emit) unless
    # null rows, which are clusters of their own, have to be spliced in
    valid, res, keys, null_rows = _merge_groups(df, group_list, sk, ek, min_dist, want_ids=True,
                                                want_row_spans=bool(return_cluster_intervals))
    n_clusters = res["n_clusters"]
    out = {}
    if null_rows is None:
        if return_cluster_ids:
            out["cluster"] = _to_host(res["ids"])
        if return_cluster_intervals:
            out["cluster_start"] = _to_host(res["row_start"])
            out["cluster_end"] = _to_host(res["row_end"])
    else:
        ids = np.full(len(df), -1)
        ids[valid] = _to_host(res["ids"])
        n_null = int(null_rows.sum())
        ids[null_rows] = n_clusters + np.arange(n_null)
        assert np.all(ids >= 0)
        if return_cluster_ids:
            out["cluster"] = ids
        if return_cluster_intervals:
            # null rows are their own "cluster" with their own (null) coordinates
            cs, ce = _to_host(res["cstart"]), _to_host(res["cend"])
            cs = pd.array(np.concatenate([cs, np.zeros(n_null, np.int64)]), dtype=pd.Int64Dtype())
            ce = pd.array(np.concatenate([ce, np.zeros(n_null, np.int64)]), dtype=pd.Int64Dtype())
            cs[n_clusters:] = df.loc[null_rows, sk].astype(pd.Int64Dtype()).values
            ce[n_clusters:] = df.loc[null_rows, ek].astype(pd.Int64Dtype()).values
            out["cluster_start"] = cs[ids]
            out["cluster_end"] = ce[ids]
    out = pd.DataFrame(out, copy=False)
    if return_input:
        out = pd.concat([df, out], axis="columns")
    if len(out) != len(df):  # nothing asked for: the reference's (discarded) set_index of the input's labels fails (ops.py:706)
        raise ValueError(f"Length mismatch: Expected {len(out)} rows, received array of length {len(df)}")
    return out


def merge(df, min_dist=0, cols=None, on=None):
    """Merge overlapping intervals (ops.py:711-839): one row per cluster with
    `n_intervals`; null rows are appended unmerged with n_intervals = NA."""
    if min_dist is not None and min_dist < 0:
        raise ValueError("min_dist>=0 currently required")
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    checks.is_bedframe(df, raise_errors=True, cols=[ck, sk, ek])
    df = df.reset_index(drop=True)  # the frame itself is only read here: no copy of its columns is needed
    group_list = _check_min_dist_and_on(df, min_dist, ck, on)

    _, res, keys, null_rows = _merge_groups(df, group_list, sk, ek, min_dist, want_ids=False)
    parts = [_cluster_table(df, res, keys, group_list, sk, ek)] if res["n_clusters"] else []
    n_null = 0 if null_rows is None else int(null_rows.sum())
    if n_null:
        extra = pd.DataFrame([pd.NA] * n_null, columns=["n_intervals"], index=df.loc[null_rows].index)
        parts.append(pd.concat([df.loc[null_rows], extra], axis=1))
    if not parts:
        raise ValueError("No objects to concatenate")  # what pd.concat([]) raises in the reference
    clusters = pd.concat(parts).reset_index(drop=True)
    if n_null:
        clusters = clusters.astype({sk: pd.Int64Dtype(), ek: pd.Int64Dtype(), "n_intervals": pd.Int64Dtype()})
    names = list(clusters.keys())
    return clusters[[ck, sk, ek] + [c for c in names if c not in (ck, sk, ek)]]


# --------------------------------------------------------------------------
# coverage
# --------------------------------------------------------------------------
def coverage(df1, df2, suffixes=("", "_"), return_input=True, cols1=None, cols2=None):
    """Base pairs of every df1 interval covered by df2 (ops.py:842-916).  The
    reference merges df2, left-overlaps and sums clipped lengths per query; the
    kernel gets the same integers from prefix sums over the merged targets
    without enumerating pairs.  Like the reference, resets df1's index in place."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    df1.reset_index(inplace=True, drop=True)
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])  # merge(df2) validates
    if len(df2) == 0:
        raise ValueError("No objects to concatenate")  # merge(df2) of an empty frame in the reference (ops.py:888)
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])  # overlap(df1, ...) validates

    codes1, keys1 = _group_keys(df1, [ck1])
    codes2, keys2 = _group_keys(df2, [ck2])
    union = sorted(set(keys1) | set(keys2), key=str)
    rank = {key: r for r, key in enumerate(union)}
    n_groups = max(len(union), 1)
    t1 = np.array([rank[k] for k in keys1], dtype=np.int32)
    t2 = np.array([rank[k] for k in keys2], dtype=np.int32)
    rows1, g1, s1, e1 = _ranked_rows(df1, codes1, t1, sk1, ek1)
    _, g2, s2, e2 = _ranked_rows(df2, codes2, t2, sk2, ek2)
    cov = np.zeros(len(df1), dtype=np.int64)
    if len(g1):
        cov = _spread(_to_host(_engine.coverage(g1, s1, e1, g2, s2, e2, n_groups)), rows1, len(df1))
    gone = _ungrouped_rows(df1, ck1, None)
    if gone is not None:
        # the reference's left join has no rows for these (`_ungrouped_rows`), and it attaches the per-row sums of
        # the others by POSITION (ops.py:903-915): they move up and the tail of the column is null.  Reproduced.
        cov = cov[~gone]
    out = pd.DataFrame({"coverage": pd.Series(cov).astype(df1[sk1].dtype)}, copy=False)
    if return_input:
        out = pd.concat([df1, out], axis="columns")
    return out


# --------------------------------------------------------------------------
# complement
# --------------------------------------------------------------------------
def complement(df, view_df=None, view_name_col="name", cols=None, cols_view=None):
    """Intervals of the view not covered by df (ops.py:1560-1687).  df intervals
    are first clipped to the view regions they overlap (the reference does that
    with `overlap(view_df, df, return_overlap=True)`, ops.py:1614-1622), then one
    kernel call computes the gaps of every region, regions in sorted name order."""
    from . import _viewframe

    ck, sk, ek = _get_default_colnames() if cols is None else cols
    _verify_columns(df, [ck, sk, ek])
    if view_df is None:
        view_df = {i: _INT64_MAX for i in set(df[ck].dropna().values)}
    ckv, skv, ekv = _get_default_colnames() if cols_view is None else cols_view
    view_df = _viewframe.make_viewframe(view_df, view_name_col=view_name_col, cols=[ckv, skv, ekv]).rename(
        columns=dict(zip([ckv, skv, ekv], [ck, sk, ek]))
    )

    names = sorted(set(view_df[view_name_col]))
    code = {name: r for r, name in enumerate(names)}
    first_row = view_df.drop_duplicates(view_name_col).set_index(view_name_col)
    bounds_lo = np.array([first_row.loc[name, sk] for name in names], dtype=np.int64)
    bounds_hi = np.array([first_row.loc[name, ek] for name in names], dtype=np.int64)
    chroms = [first_row.loc[name, ck] for name in names]

    if all(pd.api.types.is_integer_dtype(ix.dtype) for ix in (view_df.index, df.index)):
        # the clipped pieces straight from the pair list: region of every pair = name of its view row, coordinates
        # from bf_pair_coords (or the same max / min on the host for nullable columns) -- the frame `overlap` would
        # assemble around them (every column of both inputs per pair) is never read
        checks.is_bedframe(view_df, raise_errors=True, cols=[ck, sk, ek])
        checks.is_bedframe(df, raise_errors=True, cols=[ck, sk, ek])
        plain = all(_is_plain_int(f[c]) for f in (view_df, df) for c in (sk, ek))
        res = _overlap_intidxs(view_df, df, how="inner", cols1=cols, cols2=cols, _coords=plain)
        ev1, ev2 = res[0], res[1]
        if plain:
            clip_lo, clip_hi = res[2]
        else:
            clip_lo = np.maximum(_int64_coords(view_df, sk)[ev1], _int64_coords(df, sk)[ev2])
            clip_hi = np.minimum(_int64_coords(view_df, ek)[ev1], _int64_coords(df, ek)[ev2])
        row_region = np.array([code[name] for name in view_df[view_name_col]], dtype=np.int32)  # per view ROW: few
        region = row_region[ev1] if len(row_region) else np.empty(0, np.int32)
    else:  # labels that are not integers: the join decides what becomes of them (ops.overlap)
        clipped = overlap(view_df, df, return_overlap=True, how="inner", suffixes=("", "_df"), cols1=cols, cols2=cols)
        region = pd.Index(names).get_indexer(clipped[view_name_col]).astype(np.int32)
        clip_lo, clip_hi = clipped["overlap_" + sk], clipped["overlap_" + ek]

    reg, gs, ge = _engine.complement(region, np.asarray(clip_lo, dtype=np.int64), np.asarray(clip_hi, dtype=np.int64),
                                     bounds_lo, bounds_hi)
    reg = _to_host(reg).astype(np.int64)
    if len(reg):
        # one value per region, taken per gap (the reference builds both columns from Python lists of per-gap values,
        # ops.py:1672-1687: same dtypes -- the view's chromosome dtype, whatever pandas infers for the names)
        chrom_col = pd.Series(pd.Series(chroms, dtype=view_df[ck].dtype).array.take(reg))
        region_col = pd.Series(pd.Series(names).array.take(reg))
    else:
        chrom_col, region_col = pd.Series([], dtype=object), []
    return pd.DataFrame({ck: chrom_col, sk: _to_host(gs), ek: _to_host(ge), "view_region": region_col}).reset_index(drop=True)


# --------------------------------------------------------------------------
# count_overlaps / setdiff  (SURVEY.md 8f "next" #1: fused, no pair enumeration)
# --------------------------------------------------------------------------
def _overlap_counts(df1, df2, cols1, cols2, on):
    """int64 count of df2 intervals overlapping every df1 row (0 for rows with a null key)."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    by1, by2 = [ck1], [ck2]
    if on is not None:
        if not isinstance(on, list):
            raise ValueError("on=[] must be None or list")
        if (ck1 in on) or (ck2 in on):
            raise ValueError("on=[] should not contain chromosome colnames")
        _verify_columns(df1, on)
        _verify_columns(df2, on)
        by1, by2 = by1 + on, by2 + on
    codes1, keys1 = _group_keys_shared_na(df1, by1)
    codes2, keys2 = _group_keys_shared_na(df2, by2)
    # a null in an `on` column is an ordinary key shared by both frames (groupby(dropna=False), ops.py:287-288);
    # rows with a null chromosome are all-null rows and overlap nothing
    union = sorted(set(keys1) | set(keys2), key=str)
    rank = {key: r for r, key in enumerate(union)}
    t1 = np.array([rank[k] for k in keys1], dtype=np.int32)
    t2 = np.array([rank[k] for k in keys2], dtype=np.int32)
    ok1 = np.array([not _chrom_is_na(k) for k in keys1], dtype=bool)
    ok2 = np.array([not _chrom_is_na(k) for k in keys2], dtype=bool)
    rows1, g1, s1, e1 = _ranked_rows(df1, codes1, t1, sk1, ek1, keep=ok1)
    _, g2, s2, e2 = _ranked_rows(df2, codes2, t2, sk2, ek2, keep=ok2)
    counts = np.zeros(len(df1), dtype=np.int64)
    if len(g1):
        counts = _spread(_to_host(_engine.count_overlaps(g1, s1, e1, g2, s2, e2, max(len(union), 1))), rows1, len(df1))
    return counts


def count_overlaps(df1, df2, suffixes=("", "_"), return_input=True, cols1=None, cols2=None, on=None):
    """Number of df2 intervals overlapping every df1 interval (ops.py:1371-1438).  The reference left-joins
    and counts the pair list per query; the kernel counts without enumerating pairs.  Like the reference,
    resets df1's index in place."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    df1.reset_index(inplace=True, drop=True)
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])
    # the reference counts a nullable Int64 index column per query, which yields a nullable Int64 column
    gone = _ungrouped_rows(df1, ck1, on)
    if len(df2) == 0 and len(df1) - (0 if gone is None else int(gone.sum())) > 0:
        # the reference's left join looks up the label of "no partner" (-1) in df2's index (ops.py:471-478)
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    if not pd.api.types.is_integer_dtype(df2.index.dtype):
        # the reference casts the labels of the partners to Int64 (ops.py:471-478): its own join decides whether
        # such labels are accepted, and with which error they are not
        overlap(df1, df2, how="left", return_input=False, keep_order=True, suffixes=suffixes, return_index=True, on=on,
                cols1=cols1, cols2=cols2)
    counts = _overlap_counts(df1, df2, cols1, cols2, on)
    if gone is not None:  # as in coverage: the reference counts the rows its join has and attaches them by position
        counts = counts[~gone]
    out = pd.DataFrame({"count": pd.array(counts, dtype=pd.Int64Dtype())})
    if return_input:
        out = pd.concat([df1, out], axis="columns")
    return out


def setdiff(df1, df2, cols1=None, cols2=None, on=None):
    """Rows of df1 that overlap no interval of df2 (ops.py:1333-1368), original index kept."""
    return df1.iloc[np.flatnonzero(_overlap_counts(df1, df2, cols1, cols2, on) == 0)]


# --------------------------------------------------------------------------
# subtract  (SURVEY.md 8f "next" #1: fused -- merged targets + two probes per query, no pair list)
# --------------------------------------------------------------------------
def _subtract_composed(df1, df2, return_index, suffixes, cols1, cols2):
    """ops.subtract exactly as the reference composes it (ops.py:1284-1330): left overlap of df1 with
    complement(df2).  Used when df1's index does not order like its rows (the result is sorted by index
    labels, ops.py:549-550)."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    name_updates = {ck1 + suffixes[0]: ck1, "overlap_" + sk1: sk1, "overlap_" + ek1: ek1}
    for c in df1.columns:
        if c not in (ck1, sk1, ek1):
            name_updates[c + suffixes[0]] = c
    if return_index:
        name_updates["index" + suffixes[0]] = "index" + suffixes[0]
        name_updates["index" + suffixes[1]] = "complement_index" + suffixes[1]
    all_chroms = np.unique(list(pd.unique(df1[ck1].dropna())) + list(pd.unique(df2[ck2].dropna())))
    if len(all_chroms) == 0:
        raise ValueError("No chromosomes remain after dropping nulls")
    gaps = complement(df2, view_df={c: _INT64_MAX for c in all_chroms}, cols=cols2).astype(
        {sk2: pd.Int64Dtype(), ek2: pd.Int64Dtype()}
    )
    out = overlap(df1, gaps, how="left", suffixes=suffixes, return_index=return_index, return_overlap=True,
                  keep_order=True, cols1=cols1, cols2=cols2)[list(name_updates)]
    out = out.rename(columns=name_updates)
    out = out.iloc[~pd.isna(out[sk1].values)]
    out.reset_index(drop=True, inplace=True)
    if return_index:
        inds = out["index" + suffixes[0]].values
        comp = out["complement_index" + suffixes[1]].copy()
        for i in np.unique(inds):
            comp[inds == i] -= comp[inds == i].min()
        out["sub_index" + suffixes[1]] = comp.copy()
        out.drop(columns=["complement_index" + suffixes[1]], inplace=True)
    return out


def subtract(df1, df2, return_index=False, suffixes=("", "_"), cols1=None, cols2=None):
    """Intervals of df1 with everything covered by df2 removed (ops.py:1243-1330): an interval comes back
    as zero, one or several pieces; index reset, coordinates as nullable Int64.  With return_index=True
    also 'index'+suffixes[0] (label of the source row) and 'sub_index'+suffixes[1] (ordinal of the piece).

    The reference left-joins df1 with complement(df2) and sorts by index labels.  When df1's index is an
    increasing integer index (e.g. the default RangeIndex) that order is df1's row order and one fused
    kernel call yields the pieces; otherwise the reference's composition runs on the device overlap."""
    ck1, sk1, ek1 = _get_default_colnames() if cols1 is None else cols1
    ck2, sk2, ek2 = _get_default_colnames() if cols2 is None else cols2
    _verify_columns(df1, [ck1, sk1, ek1])
    _verify_columns(df2, [ck2, sk2, ek2])
    ix = df1.index
    fused = pd.api.types.is_integer_dtype(ix.dtype) and ix.is_monotonic_increasing and ix.is_unique
    fused = fused and pd.api.types.is_integer_dtype(df2.index.dtype)  # other labels: the joins of the composition decide
    if not fused:
        return _subtract_composed(df1, df2, return_index, suffixes, cols1, cols2)

    codes1, keys1 = _group_keys(df1, [ck1])
    codes2, keys2 = _group_keys(df2, [ck2])
    if len(keys1) + len(keys2) == 0:
        raise ValueError("No chromosomes remain after dropping nulls")
    checks.is_bedframe(df2, raise_errors=True, cols=[ck2, sk2, ek2])  # overlap(view, df2) inside complement
    checks.is_bedframe(df1, raise_errors=True, cols=[ck1, sk1, ek1])  # overlap(df1, gaps)
    union = sorted(set(keys1) | set(keys2), key=str)
    rank = {key: r for r, key in enumerate(union)}
    t1 = np.array([rank[k] for k in keys1], dtype=np.int32)
    t2 = np.array([rank[k] for k in keys2], dtype=np.int32)
    rows1, g1, s1, e1 = _ranked_rows(df1, codes1, t1, sk1, ek1)
    _, g2, s2, e2 = _ranked_rows(df2, codes2, t2, sk2, ek2)
    row = ps = pe = sub = np.empty(0, dtype=np.int64)
    row_dev = None
    if len(g1):
        row_dev, ps, pe, sub = _engine.subtract(g1, s1, e1, g2, s2, e2, max(len(union), 1),
                                                want_sub_index=bool(return_index))
        row, ps, pe = _to_host(row_dev), _to_host(ps), _to_host(pe)
        if rows1 is not None:
            row, row_dev = rows1[row], None
    # the pieces' coordinates replace the source rows' own: only the other columns are taken per piece
    others = [c for c in df1.columns if c not in (ck1, sk1, ek1)]
    if df1.columns.is_unique:
        out = _taken_rows(df1[[ck1, *others]], row, row_dev, "")
        out.insert(1, sk1, _int64_array(ps))
        out.insert(2, ek1, _int64_array(pe))
    else:
        out = df1.iloc[row][[ck1, sk1, ek1, *others]].reset_index(drop=True)
        out[sk1] = _int64_array(ps)
        out[ek1] = _int64_array(pe)
    del row_dev
    if return_index:
        out["index" + suffixes[0]] = pd.array(np.asarray(ix)[row], dtype=pd.Int64Dtype())
        out["sub_index" + suffixes[1]] = pd.array(_to_host(sub) if len(row) else sub, dtype=pd.Int64Dtype())
    return out


# --------------------------------------------------------------------------
# assign_view / trim  (SURVEY.md 8f "next" #1): frame glue over the device overlap
# --------------------------------------------------------------------------
def assign_view(df, view_df, drop_unassigned=False, df_view_col="view_region", view_name_col="name", cols=None,
                cols_view=None):
    """Name of the view region every interval overlaps most (ops.py:1807-1901).  The pair list and the
    clipped coordinates come from the device (left overlap + bf_pair_coords); picking the longest overlap
    per row is the reference's sort / drop_duplicates on the assembled frame.  Resets the index."""
    from . import _viewframe

    _ck, sk, ek = _get_default_colnames() if cols is None else cols
    df = df.copy()
    df.reset_index(inplace=True, drop=True)
    checks.is_bedframe(df, raise_errors=True, cols=cols)
    view_df = _viewframe.make_viewframe(view_df, view_name_col=view_name_col, cols=cols_view)
    hits = overlap(df, view_df, how="left", suffixes=("", "_view"), return_overlap=True, keep_order=False,
                   return_index=True, cols1=cols, cols2=cols_view)
    hits["overlap_length"] = hits["overlap_" + ek] - hits["overlap_" + sk]
    best = hits.sort_values("overlap_length", ascending=False).drop_duplicates("index", keep="first").sort_values("index")
    best.rename(columns={view_name_col + "_view": df_view_col}, inplace=True)
    if drop_unassigned:
        best = best.iloc[pd.isna(best).any(axis=1).values == 0, :]
    best.reset_index(inplace=True, drop=True)
    return best[[*df.columns, df_view_col]]


def trim(df, view_df=None, df_view_col=None, view_name_col="name", return_view_columns=False, cols=None,
         cols_view=None):
    """Clip every interval to its view region; intervals outside every region become null (ops.py:1441-1557).
    Without a view the intervals are only cut at zero.  Regions come from `df_view_col` when given, else
    from `assign_view` (device overlap)."""
    from . import _viewframe

    ck, sk, ek = _get_default_colnames() if cols is None else cols
    _verify_columns(df, [ck, sk, ek])
    keep_cols = list(df.columns)
    out = df.copy()
    inferred = view_df is None
    if inferred:
        df_view_col = ck
        view_df = {c: _INT64_MAX for c in set(df[ck].copy().dropna().values)}
    ckv, skv, ekv = _get_default_colnames() if cols_view is None else cols_view
    view_df = _viewframe.make_viewframe(view_df, view_name_col=view_name_col, cols=[ckv, skv, ekv]).rename(
        columns=dict(zip([ckv, skv, ekv], [ck, sk, ek]))
    )
    if inferred:
        view_name_col = ck
    elif df_view_col is None:
        if _verify_columns(out, ["view_region"], return_as_bool=True):
            raise ValueError("column view_region already exists in input df")
        df_view_col = "view_region"
        out = assign_view(out, view_df, drop_unassigned=False, df_view_col=df_view_col, view_name_col=view_name_col,
                          cols=cols, cols_view=cols)
    else:
        _verify_columns(out, [df_view_col])
        checks.is_cataloged(out, view_df, raise_errors=True, df_view_col=df_view_col, view_name_col=view_name_col)
    out = out.merge(view_df, how="left", left_on=df_view_col, right_on=view_name_col, suffixes=("", "_view"))
    outside = pd.isnull(out[df_view_col].values)
    if outside.any():
        out.loc[outside, [ck, sk, ek]] = pd.NA
    lower, upper = out[sk + "_view"].values, out[ek + "_view"].values
    out[sk] = out[sk].clip(lower=lower, upper=upper)
    out[ek] = out[ek].clip(lower=lower, upper=upper)
    return out if return_view_columns else out[keep_cols]
