"""Run the REFERENCE's own test files against this repo's implementation (build container only:
/root/reference does not exist on the GPU box, where this test skips).  See tests/refsuite_plugin.py."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

REF_TESTS = "/root/reference/tests"
FILES = ["test_ops.py", "test_extras.py", "test_core_checks.py", "test_ops_select.py"]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not present")
def test_reference_tests_pass_on_our_ops(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests"), PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", *[os.path.join(REF_TESTS, f) for f in FILES], "-p", "no:cacheprovider",
           "-c", os.devnull, "--rootdir", str(tmp_path), "-o", "filterwarnings=default", "-p", "refsuite_plugin",
           "-q", "--no-header"]
    proc = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    tail = proc.stdout[-3000:] + proc.stderr[-2000:]
    assert "patched with bioframe_b200" in proc.stdout, tail
    m = re.search(r"(\d+) passed", proc.stdout)
    assert m and int(m.group(1)) >= 37, tail
    assert "failed" not in proc.stdout.splitlines()[-1], tail
    assert proc.returncode == 0, tail
