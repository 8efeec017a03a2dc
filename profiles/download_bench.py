"""Host download of 8.76e8 row numbers (7 GB as int64): plain pinned copy vs bf_download_rows at several chunk
sizes / thread counts.  usage (GPU box): python profiles/download_bench.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bioframe_b200 import engine  # noqa: E402

n = 875_902_758
t = torch.randint(-1, 100_000_000, (n,), dtype=torch.int64, device="cuda")
out = np.zeros(n, dtype=np.int64)
pinned = torch.empty(n, dtype=torch.int64).pin_memory()


def timed(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        w = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - w)
    return best * 1e3


print(f"pinned int64 copy: {timed(lambda: pinned.copy_(t, non_blocking=True)):.1f} ms")
for chunk in (1 << 21, 1 << 22, 1 << 23, 1 << 24):
    engine._DownloadStage.CHUNK_ROWS = chunk
    engine._DownloadStage._cache.clear()
    for threads in (4, 6, 8, 10):
        ms = timed(lambda: engine.download_rows(t, out=out, n_threads=threads))
        print(f"chunk {chunk:>9d} rows, {threads:2d} threads: {ms:.1f} ms  ({8 * n / ms / 1e6:.1f} GB/s of int64 delivered)")
assert np.array_equal(out[:: 1009], t[:: 1009].cpu().numpy())
