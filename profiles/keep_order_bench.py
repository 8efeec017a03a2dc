"""Time bf_sort_pairs (device-side keep_order) on the result of the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bioframe_b200 import engine

n1, n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
(q, t) = bench.workload(n1, n2, 0)
d = [torch.from_numpy(a).cuda() for a in (*q, *t)]
for it in range(3):
    ev1, ev2, _ = engine.overlap_intidxs(*d, bench.N_GROUPS, how="left")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    engine.sort_pairs(ev1, ev2, n1 - 1, n2 - 1)
    b.record()
    torch.cuda.synchronize()
    print(f"rows {ev1.numel()}  sort_pairs {a.elapsed_time(b):.2f} ms")
    ok = bool(((ev1[1:] > ev1[:-1]) | ((ev1[1:] == ev1[:-1]) & (ev2[1:] >= ev2[:-1]))).all().item()) if it == 2 else None
    del ev1, ev2
print("sorted:", ok)
