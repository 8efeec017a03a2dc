"""torchrun --nproc-per-node N profiles/multigpu_frames_check.py
DataFrames in -> `sharded.overlap_intidxs_frames` on N GPUs (queries sharded by (chrom, start) range, sorted targets
broadcast) -> every rank holds the whole `_overlap_intidxs` result; rank 0 compares it, chromosome block by block,
with the single-GPU glue `ops._overlap_intidxs` and with the vendored reference on chr21/chr22."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bioframe_b200 import ops, sharded, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n1, n2 = 20_000_000, 2_000_000
q, t = synth.make_intervals(n1, 1, "geom2k"), bench.targets(n2)
o = np.lexsort([q[1], q[0]])  # a frame sorted by chromosome and start, the usual state of a BED file
df1 = bench.reference_frame(*(x[o] for x in q))
df2 = bench.reference_frame(*t)
for how in ("left", "inner"):
    dist.barrier()
    w0 = time.perf_counter()
    a, b = sharded.overlap_intidxs_frames(df1, df2, how=how)
    dt = time.perf_counter() - w0
    if rank == 0:
        w1, w2 = ops._overlap_intidxs(df1, df2, how=how)
        chrom = df1["chrom"].cat.codes.to_numpy()
        ok = len(a) == len(w1)
        for c in range(24):
            mg, mw = chrom[a] == c, chrom[w1] == c
            ok = ok and np.array_equal(a[mg], w1[mw]) and np.array_equal(b[mg], w2[mw])
        ref_ok = None
        bf = bench.reference_module()
        if bf is not None:
            sub1 = df1[df1["chrom"].isin(["chr21", "chr22"])]
            sub2 = df2[df2["chrom"].isin(["chr21", "chr22"])]
            r1, r2 = bf.ops._overlap_intidxs(sub1.reset_index(drop=True), sub2.reset_index(drop=True), how=how)
            rows1, rows2 = sub1.index.to_numpy(), sub2.index.to_numpy()
            r1 = rows1[r1]
            r2 = np.where(r2 >= 0, rows2[np.maximum(r2, 0)], -1)
            ref_ok = True
            for name in ("chr21", "chr22"):
                c = synth.HG38_CHROMS.index(name)
                mg, mw = chrom[a] == c, chrom[r1] == c
                ref_ok = ref_ok and np.array_equal(a[mg], r1[mw]) and np.array_equal(b[mg], r2[mw])
        print(f"how={how} world={world}: {len(a)} rows in {dt * 1e3:.0f} ms (frames in, gathered result out); equal to the "
              f"single-GPU glue block by block: {ok}; chr21/chr22 equal to bioframe.ops._overlap_intidxs: {ref_ok}", flush=True)
dist.destroy_process_group()
