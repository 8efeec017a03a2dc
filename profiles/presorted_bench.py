"""Time overlap / cluster on inputs that are already in (chromosome, start) order (no radix passes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bioframe_b200 import engine

(q, t) = bench.workload(100_000_000, 10_000_000, 0)
def srt(x):
    o = np.lexsort([x[1], x[0]])
    return tuple(a[o] for a in x)
for name, (qq, tt) in (("unsorted", (q, t)), ("sorted", (srt(q), srt(t)))):
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (*qq, *tt)]
    for _ in range(3):
        r = engine.overlap_intidxs(*d, bench.N_GROUPS, how="left"); del r
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        r = engine.overlap_intidxs(*d, bench.N_GROUPS, how="left"); del r
    b.record(); torch.cuda.synchronize()
    print(name, "overlap left 1e8 x 1e7:", round(a.elapsed_time(b) / 3, 2), "ms")
    for _ in range(2):
        r = engine.cluster(d[0], d[1], d[2], bench.N_GROUPS); del r
    a.record()
    for _ in range(3):
        r = engine.cluster(d[0], d[1], d[2], bench.N_GROUPS); del r
    b.record(); torch.cuda.synchronize()
    print(name, "cluster 1e8:", round(a.elapsed_time(b) / 3, 2), "ms")
    for _ in range(2):
        r = engine.overlap_ordered(*d, bench.N_GROUPS); del r
    a.record()
    for _ in range(3):
        r = engine.overlap_ordered(*d, bench.N_GROUPS); del r
    b.record(); torch.cuda.synchronize()
    print(name, "left join in row order (keep_order) 1e8 x 1e7:", round(a.elapsed_time(b) / 3, 2), "ms")
    del d
