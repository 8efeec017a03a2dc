"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/bioframe_b200.h declares; the ctypes binding covers the same set.
No compute is called here (there is no GPU in the build container)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "bioframe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bioframe_b200 import _lib, build

    build.build()
    handle = _lib.lib()
    names = _declared()
    assert len(names) >= 18
    for name in names:
        assert hasattr(handle, name), name
    assert sorted(_lib.SIGNATURES) == names
    assert handle.bf_version() >= 100
    assert handle.bf_prof_num_classes() > 0


def test_workspace_queries_are_host_only():
    from bioframe_b200 import _lib

    L = _lib.lib()
    small = L.bf_overlap_workspace_bytes(1000, 1000, 24, 0)
    big = L.bf_overlap_workspace_bytes(100_000_000, 10_000_000, 24, 1)
    assert 0 < small < big < 40 * 2**30
    assert L.bf_cluster_workspace_bytes(100_000_000, 24) < 12 * 2**30
    assert L.bf_overlap_workspace_bytes(-1, 0, 1, 0) == 0


def test_no_gpu_means_loud_failure():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bioframe_b200 import _lib, engine

    with pytest.raises(_lib.EngineError):
        engine.overlap_intidxs(None, [1], [2], None, [1], [2], 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bioframe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), f
                assert "importlib" not in src and "__import__" not in src, f
