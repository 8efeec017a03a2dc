"""cProfile of the frame assembly of bioframe_b200.overlap(how='inner') at 1e7 x 1e6 on the GPU box
(where the glue spends its time once the kernels are a millisecond): python profiles/frame_profile.py"""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bioframe_b200 as ours  # noqa: E402
from bioframe_b200 import synth  # noqa: E402

n1, n2 = 10_000_000, 1_000_000
q, t = synth.make_intervals(n1, 1, "geom1k"), synth.make_intervals(n2, 2, "geom1k")
df1, df2 = bench.reference_frame(*q), bench.reference_frame(*t)
df1["name"] = np.array([f"q{i}" for i in range(1000)], dtype=object)[np.arange(n1) % 1000]
df2["score"] = (np.arange(n2) % 97).astype(np.float64)
ours.overlap(df1, df2, how="inner")
pr = cProfile.Profile()
pr.enable()
out = ours.overlap(df1, df2, how="inner")
pr.disable()
print(len(out), out.dtypes.to_dict())
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
