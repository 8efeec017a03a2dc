"""engine.download_rows into a FRESH host array (what the frame glue does) vs a pre-touched one (what bench.py's e2e
leg does): the first-touch page faults of the fresh array.  Measured on the B200 box (THP mode "madvise"): pre-touched
107 ms, fresh 239 ms per 8.76e8-row array; asking for huge pages with madvise(MADV_HUGEPAGE) on the 2 MB-aligned interior
changed nothing (240 ms), so the engine does not do it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bioframe_b200 import engine

n = 875_902_758
t = torch.randint(0, 100_000_000, (n,), dtype=torch.int64, device="cuda")
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
pre = np.zeros(n, dtype=np.int64)
engine.download_rows(t, out=pre)
for name, fn in (("pre-touched", lambda: engine.download_rows(t, out=pre)),
                 ("fresh np.empty (4 KB pages)", lambda: engine.download_rows(t))):
    best = 1e9
    for _ in range(3):
        w0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - w0); del r
    print(f"{name}: {best * 1e3:.0f} ms")
