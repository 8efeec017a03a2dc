"""The fuzz, edge, sharding and stress cases once more on the RANGE-CHECKED build of the library
(`python -m bioframe_b200.build --checked`: the same sources with -DBF_CHECKED, where every index a kernel computes
for a scattered global load or store -- radix scatter destinations, finishing-pass moves, search bounds, emit
partner and output rows -- is checked on the device and the kernel traps on a violation).  compute-sanitizer is
closed on the GPU pool (profiles/NOTES.md); this is the memory-safety evidence the repo can produce itself: the
results stay bit-exact with the oracle and no check fires."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

FILES = ["test_gpu_fuzz.py", "test_gpu_edge.py", "test_gpu_prepared.py", "test_gpu_stress.py", "test_gpu_overlap.py",
         "test_gpu_cluster.py", "test_gpu_closest.py"]


def test_suites_pass_on_the_range_checked_build(tmp_path):
    lib = os.path.join(ROOT, "bioframe_b200", "libbioframe_b200_checked.so")
    if not os.path.exists(lib):
        from bioframe_b200 import build

        build.build(checked=True)
    env = dict(os.environ, BIOFRAME_B200_LIB=lib, PYTHONDONTWRITEBYTECODE="1")
    which = subprocess.run([sys.executable, "-c", "from bioframe_b200 import _lib; print(_lib.lib().bf_checked_build())"],
                           cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert which.stdout.strip().endswith("1"), which.stdout + which.stderr  # the subprocesses really load the checked build
    cmd = [sys.executable, "-m", "pytest", *[os.path.join(ROOT, "tests", f) for f in FILES], "-m", "gpu", "-q", "-x",
           "-p", "no:cacheprovider", "--no-header"]
    proc = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = proc.stdout[-2500:] + proc.stderr[-1500:]
    assert proc.returncode == 0, tail
    assert "BF_DEVICE_CHECK failed" not in proc.stdout + proc.stderr, tail
    assert " passed" in proc.stdout, tail
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, "checked_build.log"), "w") as f:
            f.write(proc.stdout[-4000:])
