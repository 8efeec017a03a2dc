"""Build libbioframe_b200.so in-tree with nvcc for sm_100a (no torch needed).

    python -m bioframe_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB = os.path.join(PKG, "libbioframe_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-extended-lambda", "--expt-relaxed-constexpr",
]
OBJ_DIR = os.path.join(PKG, "build")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    return _stale(LIB)


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        log = (proc.stdout + proc.stderr).strip().splitlines()
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + "\n".join(log[:60]))
    return proc.stderr


LIB_CHECKED = os.path.join(PKG, "libbioframe_b200_checked.so")


def build(force=False, verbose=False, checked=False):
    """One nvcc -c per translation unit (in parallel), then one link.
    checked=True: the same sources with -DBF_CHECKED (device-side range checks that trap, common.cuh) into
    libbioframe_b200_checked.so; loaded instead of the product library when BIOFRAME_B200_LIB points to it."""
    if checked:
        return _build(LIB_CHECKED, os.path.join(PKG, "build_checked"), ["-DBF_CHECKED"], force, verbose)
    return _build(LIB, OBJ_DIR, [], force, verbose)


def _build(LIB, OBJ_DIR, defines, force, verbose):
    if not force and os.path.exists(LIB) and not _stale(LIB):
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = (["-Xptxas=-v"] if verbose else []) + list(defines)
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        jobs.append((obj, [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-c", src, "-o", obj]))
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        logs = list(pool.map(lambda j: _run(j[1]), jobs))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *[j[0] for j in jobs]])
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
