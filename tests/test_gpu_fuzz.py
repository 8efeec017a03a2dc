"""A short, deterministic slice of tests/fuzz_gpu.py (every entry point against the oracle on random
configurations: ties, points, negative coordinates, INT64_MAX ends, many groups, all join kinds)."""
import pytest

pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 4])
def test_fuzz_rounds(seed):
    import fuzz_gpu

    rounds = fuzz_gpu.run(budget=600.0, seed=seed, max_rounds=25, sizes1=(0, 1, 7, 300, 2500, 6000),
                          sizes2=(0, 1, 5, 200, 3000, 5000))
    assert rounds == 25
