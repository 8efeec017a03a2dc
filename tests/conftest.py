import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


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
