"""Interval text -> device columns (SURVEY.md 8f "next" #4): the device-side counterpart of
`bioframe.read_table(path, schema="bed3" ...)` (reference src/bioframe/io/fileops.py:42-84, i.e.
`pandas.read_csv(sep="\\t", header=None)`) for the three columns the interval operations consume.  The file
is copied to the GPU as bytes; line splitting, integer parsing and the dictionary encoding of the chromosome
names happen there (csrc/ingest.cu), so that neither a pandas frame nor the host-side groupby/encode step
(ops.py:287-289) is needed before `engine.overlap_intidxs` and friends."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import pandas as pd
import torch

from . import _lib
from .engine import Workspace, _ptr, _require_cuda, _stream_ptr

__all__ = ["DeviceIntervals", "read_bed_device"]

_MAX_NAMES = 1 << 15
_SLOTS = 1 << 16


@dataclass
class DeviceIntervals:
    """chrom: int32 code per row = position of the row's name in `names` (sorted, like pd.factorize(sort=True));
    start / end: int64.  All three on the device, rows in file order."""

    names: list
    chrom: torch.Tensor
    start: torch.Tensor
    end: torch.Tensor

    def __len__(self):
        return int(self.start.numel())

    def to_frame(self, cols=("chrom", "start", "end")):
        """The frame `read_table(path, schema="bed3")` gives for these columns."""
        names = np.array(self.names, dtype=object)
        codes = self.chrom.cpu().numpy()
        chrom = names[codes] if len(codes) else np.empty(0, dtype=object)
        return pd.DataFrame({cols[0]: chrom, cols[1]: self.start.cpu().numpy(), cols[2]: self.end.cpu().numpy()})


def read_bed_device(source, device=None):
    """Parse tab-separated interval text on the GPU.  `source`: a path, bytes, or a uint8 numpy array holding
    lines `<chrom> TAB <start> TAB <end> [TAB more columns]` ("\\n" or "\\r\\n" endings, blank lines skipped as
    pandas does).  Malformed lines raise ValueError naming the line."""
    _require_cuda()
    L = _lib.lib()
    device = torch.device(device if device is not None else "cuda")
    if isinstance(source, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(source, dtype=np.uint8)
    elif isinstance(source, np.ndarray):
        raw = np.ascontiguousarray(source, dtype=np.uint8)
    else:
        raw = np.fromfile(source, dtype=np.uint8)
    n_bytes = int(raw.size)
    empty = DeviceIntervals([], torch.empty(0, dtype=torch.int32, device=device),
                            torch.empty(0, dtype=torch.int64, device=device), torch.empty(0, dtype=torch.int64, device=device))
    if n_bytes == 0:
        return empty
    with torch.cuda.device(device):
        text = torch.from_numpy(raw.copy() if not raw.flags.writeable else raw).to(device)
        stream = _stream_ptr(device)
        ws = Workspace.get(device, max(L.bf_bed_scan_workspace_bytes(n_bytes), 256))
        n_lines = C.c_int64(0)
        _lib.check(L.bf_bed_scan(_ptr(text), n_bytes, _ptr(ws), ws.numel(), stream, C.byref(n_lines)), "bf_bed_scan")
        if n_lines.value == 0:
            return empty
        ws = Workspace.get(device, max(L.bf_bed_parse_workspace_bytes(n_bytes, n_lines.value), 256))
        plan = _lib.Plan()
        n_rows, n_names = C.c_int64(0), C.c_int32(0)
        first = np.zeros(_MAX_NAMES, dtype=np.int64)
        length = np.zeros(_MAX_NAMES, dtype=np.int32)
        slot = np.zeros(_MAX_NAMES, dtype=np.int32)
        rc = L.bf_bed_parse(_ptr(text), n_bytes, n_lines.value, _ptr(ws), ws.numel(), C.byref(plan), stream,
                            C.byref(n_rows), C.byref(n_names), C.c_void_p(first.ctypes.data),
                            C.c_void_p(length.ctypes.data), C.c_void_p(slot.ctypes.data))
        _lib.check(rc, "bf_bed_parse")
        k = n_names.value
        found = [bytes(raw[first[i]: first[i] + length[i]]).decode() for i in range(k)]
        order = sorted(range(k), key=lambda i: found[i])
        slot_code = np.full(_SLOTS, -1, dtype=np.int32)
        for code, i in enumerate(order):
            slot_code[slot[i]] = code
        chrom = torch.empty(n_rows.value, dtype=torch.int32, device=device)
        start = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        end = torch.empty(n_rows.value, dtype=torch.int64, device=device)
        rc = L.bf_bed_emit(C.byref(plan), _ptr(ws), C.c_void_p(slot_code.ctypes.data), _ptr(chrom), _ptr(start), _ptr(end),
                           stream)
        _lib.check(rc, "bf_bed_emit")
    return DeviceIntervals([found[i] for i in order], chrom, start, end)
