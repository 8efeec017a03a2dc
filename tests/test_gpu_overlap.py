"""CUDA overlap path against the CPU oracle, through the C ABI (engine.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ops_oracle as oo  # noqa: E402


def _oracle_frame(g1, s1, e1, g2, s2, e2, nb, how):
    return oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how)


def _rand(rng, n, nb, span, maxlen, p_point=0.1, neg=False):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(-span if neg else 0, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return g, s, s + ln


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
@pytest.mark.parametrize("n1,n2,nb,span", [(0, 0, 1, 10), (5, 0, 3, 50), (0, 7, 3, 50), (1, 1, 1, 5),
                                            (300, 200, 1, 500), (5000, 3000, 7, 4000),
                                            (20000, 9000, 25, 100000), (9000, 30000, 300, 3000)])
def test_overlap_matches_oracle(how, n1, n2, nb, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 * 7 + n2 * 3 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, 60, neg=(nb == 7))
    g2, s2, e2 = _rand(rng, n2, nb, span, 60, neg=(nb == 7))
    ev1, ev2, n_pairs = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, nb, how=how)
    o1, o2 = _oracle_frame(g1, s1, e1, g2, s2, e2, nb, how)
    assert ev1.shape[0] == o1.shape[0]
    assert np.array_equal(ev1.cpu().numpy(), o1)
    assert np.array_equal(ev2.cpu().numpy(), o2)
    assert n_pairs == int(((o1 >= 0) & (o2 >= 0)).sum())


def test_overlap_closed_and_golden(arrops_golden):
    from bioframe_b200 import engine

    g = arrops_golden
    for name in g["names"]:
        if name == "wide":
            continue  # needs the wide-coordinate path
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        for closed in (0, 1):
            ev1, ev2, _ = engine.overlap_intidxs(None, s1, e1, None, s2, e2, 1, how="inner", closed=bool(closed))
            assert np.array_equal(ev1.cpu().numpy(), g[f"{name}/ov{closed}_1"])
            assert np.array_equal(ev2.cpu().numpy(), g[f"{name}/ov{closed}_2"])


def test_overlap_medium_synthetic():
    """cfg0-shaped (scaled down) hg38 synthetic, all four joins."""
    from bioframe_b200 import engine, synth

    g1, s1, e1 = synth.make_intervals(200_000, 1)
    g2, s2, e2 = synth.make_intervals(150_000, 2)
    s1, e1, s2, e2 = s1 // 50, e1 // 50 + 1, s2 // 50, e2 // 50 + 1
    for how in ("inner", "outer"):
        ev1, ev2, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 24, how=how)
        o1, o2 = _oracle_frame(g1, s1, e1, g2, s2, e2, 24, how)
        assert np.array_equal(ev1.cpu().numpy(), o1)
        assert np.array_equal(ev2.cpu().numpy(), o2)
