"""Run the REFERENCE's own test files against this repo's implementation (tests/refsuite_plugin.py).

Two runs of the same files: on the oracle-backed engine double (CPU, build container: checks the pandas glue)
and -- marked gpu -- on the CUDA engine (GPU box, where the reference and its tests are the git-ignored copy
vendored by oracle/make_ref.py).  The log of the CUDA run is kept under gpurun_out/ for profiles/."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda, reference_paths

FILES = ["test_ops.py", "test_extras.py", "test_core_checks.py", "test_ops_select.py"]
REF_SRC, REF_TESTS = reference_paths()
needs_ref = pytest.mark.skipif(REF_TESTS is None, reason="reference checkout / vendored copy not present")


def _run(tmp_path, engine):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests"), PYTHONDONTWRITEBYTECODE="1",
               REFSUITE_ENGINE=engine)
    cmd = [sys.executable, "-m", "pytest", *[os.path.join(REF_TESTS, f) for f in FILES], "-p", "no:cacheprovider",
           "-c", os.devnull, "--rootdir", str(tmp_path), "-o", "filterwarnings=default", "-p", "refsuite_plugin",
           "-q", "--no-header", "-rs"]
    proc = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    tail = proc.stdout[-3000:] + proc.stderr[-2000:]
    assert "patched with bioframe_b200" in proc.stdout, tail
    m = re.search(r"(\d+) passed", proc.stdout)
    assert m and int(m.group(1)) >= 37, tail
    assert "failed" not in proc.stdout.splitlines()[-1], tail
    assert proc.returncode == 0, tail
    return proc.stdout


@needs_ref
def test_reference_tests_pass_on_our_ops(tmp_path):
    out = _run(tmp_path, "oracle")
    assert "oracle test double" in out


@needs_ref
@pytest.mark.gpu
def test_reference_tests_pass_on_cuda_engine(tmp_path):
    assert has_cuda()
    out = _run(tmp_path, "cuda")
    assert "CUDA engine" in out, out[-2000:]
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, "refsuite_cuda.log"), "w") as f:
            f.write(out)
