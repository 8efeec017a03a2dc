"""Where do the device-side group encoding and column gathers of the frame glue start to pay?  Times
`bioframe_b200.overlap(df1, df2, how='inner')` (categorical chrom + a string column | a float column) at a few sizes
with the thresholds `ops._DEVICE_ENCODE_MIN_ROWS` / `ops._DEVICE_GATHER_MIN_ROWS` forced to 0 (device) and to a
huge value (host), best of 5.  Run on the GPU box: python profiles/threshold_sweep.py"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import bioframe_b200 as ours  # noqa: E402
from bioframe_b200 import ops, synth  # noqa: E402


def best(fn, reps=5):
    fn()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return 1e3 * min(out)


SIZES = ((100_000, 100_000), (300_000, 300_000), (1_000_000, 1_000_000), (3_000_000, 300_000), (10_000_000, 1_000_000))


def main():
    huge = 1 << 62
    for n1, n2 in SIZES:
        q, t = synth.make_intervals(n1, 1, "geom1k"), synth.make_intervals(n2, 2, "geom1k")
        df1, df2 = bench.reference_frame(*q), bench.reference_frame(*t)
        df1["name"] = np.array([f"q{i}" for i in range(1000)], dtype=object)[np.arange(n1) % 1000]
        df2["score"] = (np.arange(n2) % 97).astype(np.float64)
        row = {}
        for label, enc, gat in (("host", huge, huge), ("dev_encode", 0, huge), ("dev_gather", huge, 0), ("dev_both", 0, 0)):
            ops._DEVICE_ENCODE_MIN_ROWS, ops._DEVICE_GATHER_MIN_ROWS = enc, gat
            row[label] = round(best(lambda: ours.overlap(df1, df2, how="inner")), 2)
            with ops.profile() as prof:
                out = ours.overlap(df1, df2, how="inner")
            row[label + "_stages"] = {k: round(1e3 * v, 2) for k, v in prof.items()}
        for threads in (1, 8):
            ops._ARROW_TAKE_THREADS = threads
            row[f"dev_both_take{threads}"] = round(best(lambda: ours.overlap(df1, df2, how="inner")), 2)
        print(n1, n2, len(out), row, flush=True)


if __name__ == "__main__":
    main()
