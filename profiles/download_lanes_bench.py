"""bf_download_rows with and without its second (direct int64) lane, for hosts of different speed: a slow host is
imitated by giving the widening fewer threads.  8.76e8 row numbers (one column of the bench result) into a
page-locked destination; best of 3 per setting.  python profiles/download_lanes_bench.py [rows]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bioframe_b200 import engine  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 875_902_758
    t = torch.randint(-1, 100_000_000, (n,), dtype=torch.int64, device="cuda")
    t0 = time.perf_counter()
    out = engine.pinned_rows(n)
    out[:] = 0
    print(f"page-locked destination of {8 * n / 1e9:.1f} GB: {time.perf_counter() - t0:.1f} s", flush=True)
    plain = np.zeros(n, dtype=np.int64)
    check = None
    for threads in (8, 4, 2, 1):
        for mode, dst in (("0", out), (None, out), ("0", plain)):
            if mode is None:
                os.environ.pop("BF_DOWNLOAD_DIRECT", None)
            else:
                os.environ["BF_DOWNLOAD_DIRECT"] = mode
            best, direct = 1e9, 0
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                engine.download_rows(t, out=dst, n_threads=threads)
                best = min(best, time.perf_counter() - t0)
                direct = engine.download_direct_rows()
            if check is None:
                check = t.cpu().numpy()
            ok = bool(np.array_equal(dst, check))
            kind = "pinned" if dst is out else "pageable"
            print(f"threads={threads} lane2={'off' if mode == '0' else 'adaptive'} dst={kind}: {1e3 * best:.1f} ms, "
                  f"direct rows {direct} ({100.0 * direct / n:.0f} %), equal={ok}", flush=True)


if __name__ == "__main__":
    main()
