import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def reference_paths():
    """(package parent dir, tests dir) of the UNMODIFIED reference, or (None, None).
    Build container: /root/reference.  GPU box: the git-ignored copy oracle/make_ref.py vendored
    (baseline/_ref + oracle/_ref/tests), which travels with the gpurun snapshot."""
    if os.path.isdir("/root/reference/src/bioframe") and not os.environ.get("BF_FORCE_VENDORED_REF"):
        return "/root/reference/src", "/root/reference/tests"
    pkg, tests = os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref", "tests")
    if os.path.isdir(os.path.join(pkg, "bioframe")):
        return pkg, (tests if os.path.isdir(tests) else None)
    return None, None


def import_reference():
    """The reference `bioframe` module (a stub matplotlib goes on the path when the real one is missing)."""
    src, _ = reference_paths()
    if src is None:
        raise ImportError("reference not available (no /root/reference, no baseline/_ref)")
    sys.dont_write_bytecode = True
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        import tempfile

        stub = tempfile.mkdtemp(prefix="mplstub")
        os.makedirs(os.path.join(stub, "matplotlib"))
        for name, body in (("__init__", ""), ("pyplot", ""), ("colors", "def to_rgb(c):\n    return (0.0, 0.0, 0.0)\n")):
            with open(os.path.join(stub, "matplotlib", name + ".py"), "w") as f:
                f.write(body)
        sys.path.insert(0, stub)
    if src not in sys.path:
        sys.path.insert(0, src)
    import bioframe

    return bioframe


def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def arrops_golden():
    return np.load(os.path.join(GOLDEN, "arrops_cases.npz"))


@pytest.fixture(scope="session")
def frame_golden():
    return np.load(os.path.join(GOLDEN, "frame_cases.npz"))


def golden_frames(g):
    import pandas as pd

    out = {}
    for tag in ("df1", "df2", "df2u"):
        out[tag] = pd.DataFrame(
            {
                "chrom": g[f"{tag}/chrom"].astype(object),
                "start": g[f"{tag}/start"],
                "end": g[f"{tag}/end"],
                "strand": g[f"{tag}/strand"].astype(object),
            }
        )
    return out
