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


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() > 0 else 0)


def _as(t, dtype, device):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


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


def overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how="inner", closed=False, device=None):
    """All groups of `_overlap_intidxs` in one pass. Returns (events1, events2, n_pairs)."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    s1, e1 = _as(start1, torch.int64, device), _as(end1, torch.int64, device)
    s2, e2 = _as(start2, torch.int64, device), _as(end2, torch.int64, device)
    g1 = _as(grp1, torch.int32, device) if n_groups > 1 else None
    g2 = _as(grp2, torch.int32, device) if n_groups > 1 else None
    n1, n2 = int(s1.numel()), int(s2.numel())
    how_code = _lib.HOW[how]
    with torch.cuda.device(device):
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
        ev1 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        ev2 = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_overlap_emit(C.byref(plan), _ptr(ws), _ptr(ev1), _ptr(ev2), stream)
        _lib.check(rc, "bf_overlap_emit")
    return ev1, ev2, n_pairs.value
