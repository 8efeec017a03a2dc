"""Device-side entry points: torch CUDA tensors in, torch CUDA tensors out,
through the C ABI (ctypes).  torch is used for device memory and streams only.

Array conventions (all on one CUDA device):
  grp   int32  group rank of each row in the reference's group visiting order
  start int64, end int64
Outputs are int64 tensors of global row positions (-1 = no partner), exactly
what `bioframe.ops._overlap_intidxs` (ops.py:242-338) returns.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.EngineError("bioframe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


# The scratch of every op -- Workspace, the prepared-set stores, the download staging -- is ONE grow-only buffer
# per device that carries state from a _count call to its _emit call.  ctypes releases the GIL inside the library,
# so two Python threads on one device would overwrite each other's scratch: every op holds the device's lock
# from its first library call to its last (calls on different devices run side by side).
import contextlib
import threading

_DEVICE_LOCKS = {}
_LOCKS_LOCK = threading.Lock()


@contextlib.contextmanager
def _on(device):
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    with _LOCKS_LOCK:
        lock = _DEVICE_LOCKS.setdefault(idx, threading.RLock())
    with lock, torch.cuda.device(device):
        yield


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() > 0 else 0)


def _as(t, dtype, device):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


class _DownloadStage:
    """Device + pinned staging of `download_rows`, one per device, kept across calls."""

    CHUNK_ROWS = 1 << 23  # 32 MB of int32 per chunk (profiles/download_bench.py: best with 8 threads)
    _cache = {}

    @classmethod
    def get(cls, device):
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        if key not in cls._cache:
            nbytes = _lib.lib().bf_download_stage_bytes(cls.CHUNK_ROWS)
            dev = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            pinned = torch.empty(8 * cls.CHUNK_ROWS, dtype=torch.uint8).pin_memory()
            cls._cache[key] = (dev, pinned)
        return cls._cache[key]


def host_threads():
    """Host threads a download may use: the cores this process may run on, minus a little headroom."""
    import os

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(8, n - 2))  # the widening is bound by host memory bandwidth well before 8 threads


def pinned_rows(n):
    """A page-locked int64 host array of n values (a numpy view of a pinned torch tensor), for results that are
    downloaded again and again into the same buffer: `download_rows(t, out=...)` into such an array may use its
    second, direct lane on hosts whose threads cannot keep up with the bus."""
    return torch.empty(int(n), dtype=torch.int64, pin_memory=True).numpy()


def download_direct_rows():
    """Rows of this thread's last `download_rows` call that crossed the bus as plain int64 copies (8 instead of 4
    bytes per row) because the host threads were the slower side."""
    return int(_lib.lib().bf_download_direct_rows())


def download_rows(t, out=None, n_threads=None):
    """int64 CUDA tensor of row numbers (values in [-1, 2^31 - 1)) -> int64 numpy array on the host, crossing PCIe
    as 4 bytes per value (bf_download_rows).  `out`: optional preallocated contiguous int64 numpy array; when it
    is page-locked (`pinned_rows`) the library may send the tail of a large result straight into it."""
    import numpy as np

    _require_cuda()
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("download_rows: contiguous int64 CUDA tensor expected")
    n = int(t.numel())
    if out is None:
        out = np.empty(n, dtype=np.int64)
    if out.dtype != np.int64 or not out.flags.c_contiguous or out.size < n:
        raise ValueError("download_rows: out must be a contiguous int64 array of at least t.numel() values")
    if n == 0:
        return out[:0]
    device = t.device
    with _on(device):
        dev, pinned = _DownloadStage.get(device)
        rc = _lib.lib().bf_download_rows(
            _ptr(t), n, C.c_void_p(out.ctypes.data), _ptr(dev), dev.numel(), C.c_void_p(pinned.data_ptr()),
            pinned.numel(), int(n_threads or host_threads()), _stream_ptr(device),
        )
        _lib.check(rc, "bf_download_rows")
    return out[:n]


def to_device(x, dtype="int64", device=None):
    """numpy / tensor -> CUDA tensor (the frame glue moves a column once and hands it to several calls)."""
    _require_cuda()
    return _as(x, getattr(torch, dtype), torch.device(device if device is not None else "cuda"))


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Workspace:
    """Grow-only device scratch reused across calls (one per device)."""

    _cache = {}

    @classmethod
    def get(cls, device, nbytes):
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            cls._cache.pop(key, None)
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf

    @classmethod
    def release(cls):
        cls._cache.clear()


# ---- group batches: the wide path for key layouts that do not fit 64 bits ------------------------------------
# Every call lays all groups end to end on one linear axis (csrc/keys.cuh); with very many groups -- on=[...]
# columns multiply the chromosomes -- times long intervals the (position, length) key can need more than 64
# bits and the library answers BF_ERR_RANGE (OverflowError here).  Groups never interact, so the call is then
# repeated over contiguous ranges of groups, halving a range until its layout fits; results of the ranges are
# stitched in group order, which is the order of the whole call.  Only a SINGLE group whose own coordinates do not
# fit (spread over more than ~2^63) still raises.
def _rows_of_groups(g, lo, hi):
    return torch.nonzero((g >= lo) & (g < hi)).flatten()


def _global_rows(ev, rows):
    """row numbers inside a batch -> row numbers of the whole set (-1 stays -1)"""
    if rows.numel() == 0:
        return ev
    return torch.where(ev >= 0, rows[ev.clamp(min=0)], ev)


def _over_group_ranges(run, lo, hi):
    """[run(a, b) for consecutive group ranges [a, b) covering [lo, hi)], ranges halved on OverflowError"""
    try:
        return [run(lo, hi)]
    except OverflowError:
        if hi - lo <= 1:
            raise
        mid = (lo + hi) // 2
        return _over_group_ranges(run, lo, mid) + _over_group_ranges(run, mid, hi)


def overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=False, device=None,
                    row_offset1=0, row_offset2=0):
    """`_overlap_intidxs` for all groups (see _overlap_intidxs_once); falls back to group batches when the key
    layout of all groups together does not fit."""
    try:
        return _overlap_intidxs_once(grp1, start1, end1, grp2, start2, end2, n_groups, how, closed, device,
                                     row_offset1, row_offset2)
    except OverflowError:
        if n_groups <= 1:
            raise
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1, g2 = _as(grp1, torch.int32, device), _as(grp2, torch.int32, device)

    def run(lo, hi):
        r1, r2 = _rows_of_groups(g1, lo, hi), _rows_of_groups(g2, lo, hi)
        a, b, n_pairs = _overlap_intidxs_once(g1[r1] - lo, s1[r1], e1[r1], g2[r2] - lo, s2[r2], e2[r2], hi - lo, how,
                                              closed, device, 0, 0)
        a, b = _global_rows(a, r1), _global_rows(b, r2)
        if row_offset1:
            a = torch.where(a >= 0, a + int(row_offset1), a)
        if row_offset2:
            b = torch.where(b >= 0, b + int(row_offset2), b)
        return a, b, n_pairs

    parts = _over_group_ranges(run, 0, int(n_groups))
    return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), sum(p[2] for p in parts)


def _overlap_intidxs_once(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=False, device=None,
                          row_offset1=0, row_offset2=0):
    """All groups of `_overlap_intidxs` in one pass. Returns (events1, events2, n_pairs).
    row_offset1/2 are added to the reported row numbers (query-sharded callers)."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    how_code = _lib.HOW[how]
    with _on(device):
        nbytes = L.bf_overlap_workspace_bytes(n1, n2, n_groups, how_code)
        ws = Workspace.get(device, max(nbytes, 256))
        plan = _lib.Plan()
        n_rows = C.c_int64(0)
        n_pairs = C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_overlap_count(
            _ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups, how_code,
            int(bool(closed)), _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows), C.byref(n_pairs),
        )
        _lib.check(rc, "bf_overlap_count")
        if row_offset1 or row_offset2:
            _lib.check(L.bf_overlap_set_row_offsets(C.byref(plan), int(row_offset1), int(row_offset2)),
                       "bf_overlap_set_row_offsets")
        ev1 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ev2 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_overlap_emit(C.byref(plan), _ptr(ws), _ptr(ev1), _ptr(ev2), stream)
        _lib.check(rc, "bf_overlap_emit")
    return ev1, ev2, n_pairs.value


class SetStores:
    """Grow-only device buffers for prepared sets, one per (device, slot), kept across calls: a caller that
    prepares the same two sets step after step (sharded.overlap_range_sharded) names its slots and never goes
    back to the allocator -- a buffer used on two streams would otherwise be handed back through
    record_stream and re-allocated.  The caller orders the streams itself."""

    _cache = {}

    @classmethod
    def get(cls, device, slot, nbytes):
        key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), slot)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            cls._cache.pop(key, None)
            buf = None
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf

    @classmethod
    def release(cls):
        cls._cache.clear()


class SortedSet:
    """A set sorted once on the device (bf_overlap_prepare_set) or adopted from another rank
    (bf_overlap_adopt_set): `keys` (int64 view of the uint64 sort keys) and `ids` (int32 view of the uint32 row
    ids) are the two buffers a broadcast moves; `store` owns the memory."""

    def __init__(self, struct, store, n, n_groups, grp, device):
        self.struct, self.store, self.n, self.n_groups, self.grp, self.device = struct, store, int(n), int(n_groups), grp, device
        kp, ip, nn = C.c_void_p(0), C.c_void_p(0), C.c_int64(0)
        _lib.check(_lib.lib().bf_sorted_set_buffers(C.byref(struct), C.byref(kp), C.byref(ip), C.byref(nn)),
                   "bf_sorted_set_buffers")
        base = store.data_ptr()
        if self.n:
            ko, io = kp.value - base, ip.value - base
            self.keys = store[ko:ko + 8 * self.n].view(torch.int64)
            self.ids = store[io:io + 4 * self.n].view(torch.int32)
        else:
            self.keys = torch.empty(0, dtype=torch.int64, device=device)
            self.ids = torch.empty(0, dtype=torch.int32, device=device)


def set_runmax(sorted_set):
    """Running maximum of end' of a prepared set, computed once into the set's own store (bf_sorted_set_runmax)
    and reused by every `overlap_prepared` that needs it (left / outer joins: the unpaired-row test)."""
    _require_cuda()
    with _on(sorted_set.device):
        _lib.check(_lib.lib().bf_sorted_set_runmax(C.byref(sorted_set.struct), _stream_ptr(sorted_set.device)),
                   "bf_sorted_set_runmax")
    return sorted_set


def measure_extent(grp, start, end, n_groups, device=None):
    """Extent of one set as the mergeable host record of the C ABI (bf_extent_measure).
    -> (int64 numpy words, presorted flag).  Merge records with `np.maximum` / all_reduce(MAX)."""
    import numpy as np

    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s, e = _as(start, torch.int64, device), _as(end, torch.int64, device)
    g = _as(grp, torch.int32, device) if n_groups > 1 else None
    n = int(s.numel())
    words = np.zeros(int(L.bf_extent_words(n_groups)), dtype=np.int64)
    pres = C.c_int32(0)
    with _on(device):
        ws = torch.empty(max(int(L.bf_extent_workspace_bytes(n_groups)), 256), dtype=torch.uint8, device=device)
        rc = L.bf_extent_measure(_ptr(g), _ptr(s), _ptr(e), n, n_groups, _ptr(ws), ws.numel(), _stream_ptr(device),
                                 C.c_void_p(words.ctypes.data), C.byref(pres))
        _lib.check(rc, "bf_extent_measure")
    return words, bool(pres.value)


def prepare_set(words, grp, start, end, n_groups, presorted=False, device=None, slot=None):
    """Sort one set for `overlap_prepared` with the key layout the merged extent `words` fix.
    slot: name of a persistent store (SetStores) to sort into instead of a fresh allocation; the set then is
    valid until the slot is prepared again."""
    import numpy as np

    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s, e = _as(start, torch.int64, device), _as(end, torch.int64, device)
    g = _as(grp, torch.int32, device) if n_groups > 1 else None
    n = int(s.numel())
    words = np.ascontiguousarray(words, dtype=np.int64)
    struct = _lib.SortedSetStruct()
    with _on(device):
        nbytes = max(int(L.bf_sorted_set_bytes(n, n_groups, 0)), 256)
        store = SetStores.get(device, slot, nbytes) if slot else torch.empty(nbytes, dtype=torch.uint8, device=device)
        rc = L.bf_overlap_prepare_set(C.c_void_p(words.ctypes.data), _ptr(g), _ptr(s), _ptr(e), n, n_groups,
                                      int(bool(presorted)), _ptr(store), store.numel(), _stream_ptr(device),
                                      C.byref(struct))
        _lib.check(rc, "bf_overlap_prepare_set")
    return SortedSet(struct, store, n, n_groups, g, device)


def adopt_set(words, n, n_groups, grp=None, device=None, slot=None):
    """An empty prepared set of `n` rows whose `keys` / `ids` the caller fills (e.g. by a broadcast of another
    rank's `SortedSet.keys` / `.ids`) before it is used."""
    import numpy as np

    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    words = np.ascontiguousarray(words, dtype=np.int64)
    struct = _lib.SortedSetStruct()
    with _on(device):
        nbytes = max(int(L.bf_sorted_set_bytes(int(n), n_groups, 1)), 256)
        store = SetStores.get(device, slot, nbytes) if slot else torch.empty(nbytes, dtype=torch.uint8, device=device)
        rc = L.bf_overlap_adopt_set(C.c_void_p(words.ctypes.data), int(n), n_groups, _ptr(store), store.numel(),
                                    _stream_ptr(device), C.byref(struct))
        _lib.check(rc, "bf_overlap_adopt_set")
    g = _as(grp, torch.int32, device) if (grp is not None and n_groups > 1) else None
    return SortedSet(struct, store, n, n_groups, g, device)


def overlap_prepared(set1, set2, how="inner", closed=False, own_range=None, row_offset1=0, row_offset2=0,
                     n_own1=None, halo_ids=None, want_group_counts=False):
    """`_overlap_intidxs` on two prepared sets (bf_overlap_count_prepared + bf_overlap_emit).
    own_range = (g_lo, start_lo, g_hi, start_hi): produce only the part of the result this query shard owns;
    n_own1 / halo_ids: set-1 rows from n_own1 on are halo rows with these global row numbers.
    -> (events1, events2, n_pairs[, per-group counts int64[n_groups, 4] = A, B, unpaired 1, unpaired 2])."""
    import numpy as np

    _require_cuda()
    L = _lib.lib()
    device = set1.device
    how_code = _lib.HOW[how]
    own = None if own_range is None else np.ascontiguousarray(own_range, dtype=np.int64)
    with _on(device):
        ws = Workspace.get(device, max(L.bf_overlap_prepared_workspace_bytes(set1.n, set2.n, set1.n_groups, how_code), 256))
        plan = _lib.Plan()
        n_rows, n_pairs = C.c_int64(0), C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_overlap_count_prepared(C.byref(set1.struct), C.byref(set2.struct), _ptr(set1.grp), _ptr(set2.grp),
                                         how_code, int(bool(closed)), C.c_void_p(own.ctypes.data) if own is not None else None,
                                         _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows), C.byref(n_pairs))
        _lib.check(rc, "bf_overlap_count_prepared")
        if row_offset1 or row_offset2:
            _lib.check(L.bf_overlap_set_row_offsets(C.byref(plan), int(row_offset1), int(row_offset2)),
                       "bf_overlap_set_row_offsets")
        halo = None
        if n_own1 is not None and int(n_own1) < set1.n:
            halo = _as(halo_ids, torch.int64, device)
            if int(halo.numel()) != set1.n - int(n_own1):
                raise ValueError("halo_ids must hold one global row number per halo row")
            _lib.check(L.bf_overlap_set_halo_rows(C.byref(plan), int(n_own1), _ptr(halo)), "bf_overlap_set_halo_rows")
        counts = None
        if want_group_counts:
            counts = np.zeros((set1.n_groups, 4), dtype=np.int64)
            _lib.check(L.bf_overlap_group_counts(C.byref(plan), _ptr(ws), C.c_void_p(counts.ctypes.data), stream),
                       "bf_overlap_group_counts")
        ev1 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ev2 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_overlap_emit(C.byref(plan), _ptr(ws), _ptr(ev1), _ptr(ev2), stream)
        _lib.check(rc, "bf_overlap_emit")
    if want_group_counts:
        return ev1, ev2, n_pairs.value, counts
    return ev1, ev2, n_pairs.value


def sort_pairs(ev1, ev2, max1, max2):
    """Order device-resident result pairs in place by (events1, events2), -1 last (keep_order=True of
    bioframe.overlap, ops.py:549-550, for monotonic indexes)."""
    _require_cuda()
    L = _lib.lib()
    n = int(ev1.numel())
    if n == 0:
        return ev1, ev2
    device = ev1.device
    with _on(device):
        ws = Workspace.get(device, max(L.bf_sort_pairs_workspace_bytes(n), 256))
        rc = L.bf_sort_pairs(_ptr(ev1), _ptr(ev2), n, int(max1), int(max2), _ptr(ws), ws.numel(), _stream_ptr(device))
        _lib.check(rc, "bf_sort_pairs")
    return ev1, ev2


def cluster(grp, start, end, n_groups, min_dist=0, want_ids=True, want_table=True, want_row_spans=False, device=None):
    """merge_intervals over all groups (see _cluster_once); group batches when the key layout does not fit."""
    try:
        return _cluster_once(grp, start, end, n_groups, min_dist, want_ids, want_table, want_row_spans, device)
    except OverflowError:
        if n_groups <= 1:
            raise
    device = torch.device(device if device is not None else "cuda")
    s, e, g = _as(start, torch.int64, device), _as(end, torch.int64, device), _as(grp, torch.int32, device)

    def run(lo, hi):
        rows = _rows_of_groups(g, lo, hi)
        return rows, lo, _cluster_once(g[rows] - lo, s[rows], e[rows], hi - lo, min_dist, want_ids, want_table,
                                       want_row_spans, device)

    parts = _over_group_ranges(run, 0, int(n_groups))
    n = int(s.numel())
    out = dict(ids=None, n_clusters=sum(p[2]["n_clusters"] for p in parts), cgrp=None, cstart=None, cend=None,
               ccount=None, row_start=None, row_end=None)
    for name in ("ids", "row_start", "row_end"):
        if parts[0][2][name] is not None:
            out[name] = torch.empty(n, dtype=torch.int64, device=device)
    base = 0
    for rows, lo, res in parts:  # cluster ids count on from range to range (ops.py:653-655)
        if out["ids"] is not None:
            out["ids"][rows] = res["ids"] + base
        for name in ("row_start", "row_end"):
            if out[name] is not None:
                out[name][rows] = res[name]
        base += res["n_clusters"]
    if want_table:
        out["cgrp"] = torch.cat([res["cgrp"] + lo for _, lo, res in parts])
        for name in ("cstart", "cend", "ccount"):
            out[name] = torch.cat([res[name] for _, _, res in parts])
    return out


def _cluster_once(grp, start, end, n_groups, min_dist=0, want_ids=True, want_table=True, want_row_spans=False, device=None):
    """merge_intervals over all groups (group code order = sorted key order).

    Returns dict(ids=int64[n] | None, n_clusters, cgrp=int32[C], cstart, cend,
    ccount=int64[C], row_start/row_end=int64[n] | None); tensors stay on device."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s, e = _as(start, torch.int64, device), _as(end, torch.int64, device)
    g = _as(grp, torch.int32, device) if n_groups > 1 else None
    n = int(s.numel())
    has_md = 0 if min_dist is None else 1
    md = 0 if min_dist is None else int(min_dist)
    need_ids = want_ids or want_row_spans
    with _on(device):
        ws = Workspace.get(device, max(L.bf_cluster_workspace_bytes(n, n_groups), 256))
        plan = _lib.Plan()
        nc = C.c_int64(0)
        stream = _stream_ptr(device)
        ids = torch.empty(n, dtype=torch.int64, device=device) if need_ids else None
        rc = L.bf_cluster_count(_ptr(g), _ptr(s), _ptr(e), n, n_groups, md, has_md, _ptr(ws), ws.numel(),
                                C.byref(plan), stream, _ptr(ids), C.byref(nc))
        _lib.check(rc, "bf_cluster_count")
        c = nc.value
        out = dict(ids=ids, n_clusters=c, cgrp=None, cstart=None, cend=None, ccount=None, row_start=None, row_end=None)
        if want_table:
            out["cgrp"] = torch.empty(c, dtype=torch.int32, device=device)
            out["cstart"] = torch.empty(c, dtype=torch.int64, device=device)
            out["cend"] = torch.empty(c, dtype=torch.int64, device=device)
            out["ccount"] = torch.empty(c, dtype=torch.int64, device=device)
        if want_row_spans:
            out["row_start"] = torch.empty(n, dtype=torch.int64, device=device)
            out["row_end"] = torch.empty(n, dtype=torch.int64, device=device)
        if want_table or want_row_spans:
            rc = L.bf_cluster_emit(C.byref(plan), _ptr(ws), _ptr(out["cgrp"]), _ptr(out["cstart"]), _ptr(out["cend"]),
                                   _ptr(out["ccount"]), _ptr(ids), _ptr(out["row_start"]), _ptr(out["row_end"]), stream)
            _lib.check(rc, "bf_cluster_emit")
    return out


def _per_query_batched(once, grp1, start1, end1, grp2, start2, end2, n_groups, device, **kw):
    """a per-query result (coverage, counts) over group batches: out[rows of the batch] = result of the batch"""
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1, g2 = _as(grp1, torch.int32, device), _as(grp2, torch.int32, device)
    out = torch.zeros(int(s1.numel()), dtype=torch.int64, device=device)

    def run(lo, hi):
        r1, r2 = _rows_of_groups(g1, lo, hi), _rows_of_groups(g2, lo, hi)
        if r1.numel():
            out[r1] = once(g1[r1] - lo, s1[r1], e1[r1], g2[r2] - lo, s2[r2], e2[r2], hi - lo, device=device, **kw)
        return None

    _over_group_ranges(run, 0, int(n_groups))
    return out


def coverage(grp1, start1, end1, grp2, start2, end2, n_groups, device=None):
    """Covered base pairs of every row of set 1 by the union of set 2 -> int64[n1] (device).
    The kernel does arithmetic on the ends themselves, so ends of INT64_MAX (what `complement` / `subtract` /
    `trim` put at the end of a chromosome, ops.py:1309, 1499, 1604) do not fit its key layout as they are: such
    calls are repeated with every coordinate clipped to the window in which both sets have intervals --
    outside it nothing of set 1 meets anything of set 2, so no covered base pair is lost -- and, if the groups
    together still need more than 64 key bits, over group batches."""
    try:
        return _coverage_once(grp1, start1, end1, grp2, start2, end2, n_groups, device)
    except OverflowError:
        pass
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    if s1.numel() and s2.numel():
        lo = max(int(s1.min().item()), int(s2.min().item()))
        hi = min(int(e1.max().item()), int(e2.max().item()))
        hi = max(hi, lo)
        s1, e1, s2, e2 = (x.clamp(min=lo, max=hi) for x in (s1, e1, s2, e2))
    try:
        return _coverage_once(grp1, s1, e1, grp2, s2, e2, n_groups, device)
    except OverflowError:
        if n_groups <= 1:
            raise
    return _per_query_batched(_coverage_once, grp1, s1, e1, grp2, s2, e2, n_groups, device)


def _coverage_once(grp1, start1, end1, grp2, start2, end2, n_groups, device=None):
    """Covered base pairs of every row of set 1 by the union of set 2 -> int64[n1] (device)."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    with _on(device):
        ws = Workspace.get(device, max(L.bf_coverage_workspace_bytes(n1, n2, n_groups), 256))
        cov = torch.empty(n1, dtype=torch.int64, device=device)
        rc = L.bf_coverage(_ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups,
                           _ptr(ws), ws.numel(), _stream_ptr(device), _ptr(cov))
        _lib.check(rc, "bf_coverage")
    return cov


def count_overlaps(grp1, start1, end1, grp2, start2, end2, n_groups, closed=False, device=None):
    """Number of set-2 intervals overlapping every row of set 1 -> int64[n1] (device, input row order);
    group batches when the key layout of all groups together does not fit."""
    try:
        return _count_overlaps_once(grp1, start1, end1, grp2, start2, end2, n_groups, closed, device)
    except OverflowError:
        if n_groups <= 1:
            raise
    return _per_query_batched(_count_overlaps_once, grp1, start1, end1, grp2, start2, end2, n_groups, device, closed=closed)


def _count_overlaps_once(grp1, start1, end1, grp2, start2, end2, n_groups, closed=False, device=None):
    """Number of set-2 intervals overlapping every row of set 1 -> int64[n1] (device, input row order)."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    with _on(device):
        ws = Workspace.get(device, max(L.bf_count_overlaps_workspace_bytes(n1, n2, n_groups), 256))
        counts = torch.empty(n1, dtype=torch.int64, device=device)
        rc = L.bf_count_overlaps(_ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups,
                                 int(bool(closed)), _ptr(ws), ws.numel(), _stream_ptr(device), _ptr(counts))
        _lib.check(rc, "bf_count_overlaps")
    return counts


def overlap_ordered(grp1, start1, end1, grp2, start2, end2, n_groups, device=None, max_partners=None):
    """overlap(how='left') with the rows already in (row of set 1, row of set 2) order, -1 = no partner
    (bf_overlap_ordered_*).  -> (events1, events2, largest partner count of one row).
    max_partners: give up right after the count phase -- before any output is allocated or written -- when one
    row has more partners than that (the emit orders a row's partners inside one block); -> (None, None, largest)."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    with _on(device):
        ws = Workspace.get(device, max(L.bf_overlap_ordered_workspace_bytes(n1, n2, n_groups), 256))
        plan = _lib.Plan()
        n_rows, biggest = C.c_int64(0), C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_overlap_ordered_count(_ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups,
                                        _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows), C.byref(biggest))
        _lib.check(rc, "bf_overlap_ordered_count")
        if max_partners is not None and biggest.value > int(max_partners):
            return None, None, biggest.value
        ev1 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ev2 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_overlap_ordered_emit(C.byref(plan), _ptr(ws), _ptr(ev1), _ptr(ev2), stream)
        _lib.check(rc, "bf_overlap_ordered_emit")
    return ev1, ev2, biggest.value


def gather_rows(column, idx, want_valid=False, device=None):
    """column[idx] for a 1-, 2-, 4- or 8-byte column (numpy array or tensor) and device row numbers `idx`
    (-1 -> 0) -> device tensor of the column's dtype (and the validity bytes when asked)."""
    _require_cuda()
    device = torch.device(device if device is not None else "cuda")
    col = column if isinstance(column, torch.Tensor) else torch.as_tensor(column)
    if col.element_size() not in (1, 2, 4, 8) or col.dtype in (torch.complex64,):
        raise ValueError("gather_rows: 1-, 2-, 4- or 8-byte column expected")
    col = col.to(device=device, non_blocking=True).contiguous()
    ix = _as(idx, torch.int64, device)
    n = int(ix.numel())
    with _on(device):
        out = torch.empty(n, dtype=col.dtype, device=device)
        valid = torch.empty(n, dtype=torch.uint8, device=device) if want_valid else None
        rc = _lib.lib().bf_gather_rows(_ptr(col), col.element_size(), _ptr(ix), n, _ptr(out), _ptr(valid),
                                       _stream_ptr(device))
        _lib.check(rc, "bf_gather_rows")
    return (out, valid) if want_valid else out


def codes_present(codes, n_categories, device=None):
    """Which dictionary codes of a categorical key column occur (bf_group_codes_present).
    codes: int8 / int16 numpy array or CUDA tensor (-1 = null).  -> (device codes, bool numpy [n_categories + 1]:
    slot 0 = a null occurs, slot c + 1 = category c occurs)."""
    import numpy as np

    _require_cuda()
    device = torch.device(device if device is not None else "cuda")
    t = codes if isinstance(codes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(codes))
    t = t.to(device=device, non_blocking=True).contiguous()
    with _on(device):
        present = torch.empty(int(n_categories) + 2, dtype=torch.int32, device=device)
        rc = _lib.lib().bf_group_codes_present(_ptr(t), t.element_size(), int(t.numel()), int(n_categories), _ptr(present),
                                               _stream_ptr(device))
        _lib.check(rc, "bf_group_codes_present")
    flags = present.cpu().numpy().astype(bool)
    if flags[-1]:
        raise ValueError("categorical codes outside [-1, n_categories)")
    return t, flags[:-1]


def codes_apply(codes_dev, lut, device=None):
    """group rank per row = lut[code + 1] (bf_group_codes_apply).  codes_dev: what `codes_present` returned;
    lut: int32 numpy [n_categories + 1].  -> int32 CUDA tensor."""
    import numpy as np

    _require_cuda()
    device = codes_dev.device
    lut_t = torch.from_numpy(np.ascontiguousarray(lut, dtype=np.int32)).to(device)
    with _on(device):
        out = torch.empty(int(codes_dev.numel()), dtype=torch.int32, device=device)
        rc = _lib.lib().bf_group_codes_apply(_ptr(codes_dev), codes_dev.element_size(), int(codes_dev.numel()), _ptr(lut_t),
                                             int(lut_t.numel()) - 1, _ptr(out), _stream_ptr(device))
        _lib.check(rc, "bf_group_codes_apply")
    return out


def pair_coords(ev1, ev2, start1, end1, start2, end2, want_overlap=True, want_distance=False, device=None):
    """Clipped coordinates / distances of reported pairs (return_overlap, return_distance of the frame ops).
    -> (overlap_start, overlap_end, distance) int64 device tensors (None where not wanted); rows with a -1
    partner hold 0."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    a, b = _as(ev1, torch.int64, device), _as(ev2, torch.int64, device)
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    n = int(a.numel())
    with _on(device):
        os_ = torch.empty(n, dtype=torch.int64, device=device) if want_overlap else None
        oe = torch.empty(n, dtype=torch.int64, device=device) if want_overlap else None
        dist = torch.empty(n, dtype=torch.int64, device=device) if want_distance else None
        rc = L.bf_pair_coords(_ptr(a), _ptr(b), n, _ptr(s1), _ptr(e1), _ptr(s2), _ptr(e2), _ptr(os_), _ptr(oe),
                              _ptr(dist), _stream_ptr(device))
        _lib.check(rc, "bf_pair_coords")
    return os_, oe, dist


def subtract(grp1, start1, end1, grp2, start2, end2, n_groups, want_sub_index=True, device=None):
    """ops.subtract pieces for all groups (see _subtract_once); group batches when the key layout of all groups
    together does not fit (the pieces of the batches are merged back into ascending row order)."""
    try:
        return _subtract_once(grp1, start1, end1, grp2, start2, end2, n_groups, want_sub_index, device)
    except OverflowError:
        if n_groups <= 1:
            raise
    device = torch.device(device if device is not None else "cuda")
    s1, e1, g1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device), _as(grp1, torch.int32, device)
    s2, e2, g2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device), _as(grp2, torch.int32, device)

    def run(lo, hi):
        r1, r2 = _rows_of_groups(g1, lo, hi), _rows_of_groups(g2, lo, hi)
        row, ps, pe, sub = _subtract_once(g1[r1] - lo, s1[r1], e1[r1], g2[r2] - lo, s2[r2], e2[r2], hi - lo,
                                          want_sub_index, device)
        return _global_rows(row, r1), ps, pe, sub

    parts = _over_group_ranges(run, 0, int(n_groups))
    row = torch.cat([p[0] for p in parts])
    order = torch.sort(row, stable=True).indices  # the pieces of a row stay left to right
    cat = lambda i: torch.cat([p[i] for p in parts])[order]  # noqa: E731
    return row[order], cat(1), cat(2), (cat(3) if want_sub_index else None)


def _subtract_once(grp1, start1, end1, grp2, start2, end2, n_groups, want_sub_index=True, device=None):
    """Pieces of every row of set 1 that the union of set 2 does not cover (ops.subtract, ops.py:1243-1330).
    -> (row int64[R], start int64[R], end int64[R], sub_index int64[R] or None) on device: rows of set 1 in
    ascending order, the pieces of a row left to right."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    with _on(device):
        ws = Workspace.get(device, max(L.bf_subtract_workspace_bytes(n1, n2, n_groups), 256))
        plan = _lib.Plan()
        n_rows = C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_subtract_count(_ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups,
                                 _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows))
        _lib.check(rc, "bf_subtract_count")
        row = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ps = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        pe = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        sub = torch.empty(n_rows.value, dtype=torch.int64, device=device) if want_sub_index else None
        rc = L.bf_subtract_emit(C.byref(plan), _ptr(ws), _ptr(row), _ptr(ps), _ptr(pe), _ptr(sub), stream)
        _lib.check(rc, "bf_subtract_emit")
    return row, ps, pe, sub


def complement(region, start, end, bounds_lo, bounds_hi, device=None):
    """Gaps of the (already clipped) intervals inside each view region.
    -> (region int32[G], start int64[G], end int64[G]) on device, regions in code order."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s, e = _as(start, torch.int64, device), _as(end, torch.int64, device)
    b0, b1 = _as(bounds_lo, torch.int64, device), _as(bounds_hi, torch.int64, device)
    nr = int(b0.numel())
    if nr == 0:
        z = torch.empty(0, dtype=torch.int64, device=device)
        return torch.empty(0, dtype=torch.int32, device=device), z, z.clone()
    g = _as(region, torch.int32, device) if nr > 1 else None
    n = int(s.numel())
    with _on(device):
        ws = Workspace.get(device, max(L.bf_complement_workspace_bytes(n, nr), 256))
        plan = _lib.Plan()
        ng = C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_complement_count(_ptr(g), _ptr(s), _ptr(e), n, _ptr(b0), _ptr(b1), nr, _ptr(ws), ws.numel(),
                                   C.byref(plan), stream, C.byref(ng))
        _lib.check(rc, "bf_complement_count")
        reg = torch.empty(ng.value, dtype=torch.int32, device=device)
        gs = torch.empty(ng.value, dtype=torch.int64, device=device)
        ge = torch.empty(ng.value, dtype=torch.int64, device=device)
        rc = L.bf_complement_emit(C.byref(plan), _ptr(ws), _ptr(reg), _ptr(gs), _ptr(ge), stream)
        _lib.check(rc, "bf_complement_emit")
    return reg, gs, ge


def closest_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, k=1, ignore_overlaps=False,
                    ignore_upstream=False, ignore_downstream=False, along=None, self_closest=False, device=None):
    """`_closest_intidxs` for all groups (see _closest_intidxs_once); group batches when the key layout of all
    groups together does not fit."""
    kw = dict(k=k, ignore_overlaps=ignore_overlaps, ignore_upstream=ignore_upstream, ignore_downstream=ignore_downstream,
              self_closest=self_closest)
    try:
        return _closest_intidxs_once(grp1, start1, end1, grp2, start2, end2, n_groups, along=along, device=device, **kw)
    except OverflowError:
        if n_groups <= 1:
            raise
    device = torch.device(device if device is not None else "cuda")
    s1, e1, g1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device), _as(grp1, torch.int32, device)
    if self_closest:
        s2, e2, g2 = s1, e1, g1
    else:
        s2, e2, g2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device), _as(grp2, torch.int32, device)
    al = _as(along, torch.uint8, device) if along is not None else None

    def run(lo, hi):
        r1 = _rows_of_groups(g1, lo, hi)
        r2 = r1 if self_closest else _rows_of_groups(g2, lo, hi)
        a, b, n_matched = _closest_intidxs_once(g1[r1] - lo, s1[r1], e1[r1], g2[r2] - lo, s2[r2], e2[r2], hi - lo,
                                                along=al[r1] if al is not None else None, device=device, **kw)
        return _global_rows(a, r1), _global_rows(b, r2), n_matched

    parts = _over_group_ranges(run, 0, int(n_groups))
    return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), sum(p[2] for p in parts)


def _closest_intidxs_once(grp1, start1, end1, grp2, start2, end2, n_groups, k=1, ignore_overlaps=False,
                          ignore_upstream=False, ignore_downstream=False, along=None, self_closest=False, device=None):
    """All groups of `_closest_intidxs` (ops.py:919-1040) in one pass.
    `self_closest`: set 2 is set 1 (grp2/start2/end2 are ignored).  `along`: bool per
    query, False = "-" strand.  Returns (events1, events2, n_matched_rows) on device."""
    _require_cuda()
    if k < 1:
        raise ValueError("k>=1 required")
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    if self_closest:
        s2, e2, g2 = s1, e1, g1
    else:
        s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
        g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    al = _as(along, torch.uint8, device) if along is not None else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    flags = (
        (_lib.CLOSEST_IGNORE_OVERLAPS if ignore_overlaps else 0)
        | (_lib.CLOSEST_IGNORE_UPSTREAM if ignore_upstream else 0)
        | (_lib.CLOSEST_IGNORE_DOWNSTREAM if ignore_downstream else 0)
        | (_lib.CLOSEST_SELF if self_closest else 0)
    )
    with _on(device):
        ws = Workspace.get(device, max(L.bf_closest_workspace_bytes(n1, n2, n_groups), 256))
        plan = _lib.Plan()
        n_rows, n_matched = C.c_int64(0), C.c_int64(0)
        stream = _stream_ptr(device)
        rc = L.bf_closest_count(_ptr(g1), _ptr(s1), _ptr(e1), n1, _ptr(g2), _ptr(s2), _ptr(e2), n2, n_groups, int(k),
                                flags, _ptr(al), _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows),
                                C.byref(n_matched))
        _lib.check(rc, "bf_closest_count")
        ev1 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ev2 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_closest_emit(C.byref(plan), _ptr(ws), _ptr(ev1), _ptr(ev2), stream)
        _lib.check(rc, "bf_closest_emit")
    return ev1, ev2, n_matched.value
