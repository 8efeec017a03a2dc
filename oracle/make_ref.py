#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- vendors the UNMODIFIED reference (open2c/bioframe v0.8.0) for the GPU box.

    python oracle/make_ref.py            # idempotent; run by __graft_entry__.build() when /root/reference exists

`/root/reference` does not exist on the GPU box, and the reference is pure Python (nothing to compile),
so the recipe makes it travel the way the task contract names for a reference install:

  baseline/_ref/bioframe/...   the reference package, as `pip install --target baseline/_ref /root/reference`
                               would lay it down.  That pip install is tried first; it fails in this image
                               (build backend `uv_build` is not in the offline wheelhouse), so the fallback
                               is what uv_build itself does for a pure-Python project: copy `src/bioframe`.
  baseline/_ref/matplotlib/    three-file stub: `bioframe/vis.py` imports matplotlib, which is not installed
                               here; no function on the hot path touches it.
  oracle/_ref/tests/...        the reference's own test files (+ their test_data), so that
                               tests/test_reference_suite.py can run them on the CUDA engine on the box.

Both directories are git-ignored (never part of the history, never read by the product) and NOT
gpurun-ignored.  Only tests/, __graft_entry__ and bench.py's reference/cpu_baseline legs import from them.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BASELINE_REF = os.path.join(ROOT, "baseline", "_ref")
ORACLE_REF = os.path.join(ROOT, "oracle", "_ref")

STUB = {
    "__init__.py": "",
    "pyplot.py": "",
    "colors.py": "def to_rgb(c):\n    return (0.0, 0.0, 0.0)\n",
}


def _pip_install():
    """The contract's offline install.  Returns (ok, one-line outcome)."""
    tmp = "/tmp/bioframe_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns(".git"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
           "/opt/wheelhouse", "--no-deps", "--target", BASELINE_REF, tmp]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    if proc.returncode == 0 and os.path.isdir(os.path.join(BASELINE_REF, "bioframe")):
        return True, "pip install --target baseline/_ref: ok"
    last = (proc.stderr.strip().splitlines() or ["?"])[-1]
    return False, f"pip install --target baseline/_ref failed ({last}); copied src/bioframe instead"


def make(verbose=True):
    if not os.path.isdir(REF):
        return "reference checkout not present: keeping whatever baseline/_ref already holds"
    os.makedirs(BASELINE_REF, exist_ok=True)
    pkg = os.path.join(BASELINE_REF, "bioframe")
    marker = os.path.join(BASELINE_REF, "INSTALL_OUTCOME.txt")
    if os.path.isdir(pkg) and os.path.exists(marker):
        outcome = open(marker).read().strip()
    else:
        ok, outcome = _pip_install()
        if not ok:
            shutil.rmtree(pkg, ignore_errors=True)
            shutil.copytree(os.path.join(REF, "src", "bioframe"), pkg, ignore=shutil.ignore_patterns("__pycache__"))
        with open(marker, "w") as f:
            f.write(outcome + "\n")
    stub = os.path.join(BASELINE_REF, "matplotlib")
    os.makedirs(stub, exist_ok=True)
    for name, body in STUB.items():
        with open(os.path.join(stub, name), "w") as f:
            f.write(body)
    tests = os.path.join(ORACLE_REF, "tests")
    if not os.path.isdir(tests):
        os.makedirs(ORACLE_REF, exist_ok=True)
        shutil.copytree(os.path.join(REF, "tests"), tests, ignore=shutil.ignore_patterns("__pycache__"))
    if verbose:
        print(f"[make_ref] {outcome}")
    return outcome


def import_reference():
    """-> the reference `bioframe` module from baseline/_ref (raises ImportError if it was never vendored)."""
    if not os.path.isdir(os.path.join(BASELINE_REF, "bioframe")):
        raise ImportError("baseline/_ref/bioframe missing: run `python oracle/make_ref.py` in the build container")
    sys.dont_write_bytecode = True
    if BASELINE_REF not in sys.path:
        sys.path.insert(0, BASELINE_REF)
    import bioframe

    return bioframe


if __name__ == "__main__":
    make()
