"""CUDA closest path against the CPU oracle, through the C ABI (engine.py).

The oracle (like the CUDA path) orders targets with equal end / equal start in
row order (stable); the reference's unstable argsort leaves that undefined, so
random cases may contain ties -- both sides define them the same way -- while
the golden vectors from the reference itself are tie-free."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ops_oracle as oo  # noqa: E402

VARIANTS = [
    dict(k=1),
    dict(k=2),
    dict(k=5),
    dict(k=1, ignore_overlaps=True),
    dict(k=3, ignore_overlaps=True),
    dict(k=2, ignore_upstream=True),
    dict(k=2, ignore_downstream=True),
    dict(k=2, ignore_overlaps=True, ignore_upstream=True),
    dict(k=4, ignore_upstream=True, ignore_downstream=True),
]


def _rand(rng, n, nb, span, maxlen, p_point=0.1, neg=False):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(-span if neg else 0, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return g, s, s + ln


def _check(a, b):
    assert a[0].shape[0] == b[0].shape[0]
    assert np.array_equal(a[0].cpu().numpy(), b[0])
    assert np.array_equal(a[1].cpu().numpy(), b[1])


@pytest.mark.parametrize("kw", VARIANTS)
@pytest.mark.parametrize("n1,n2,nb,span", [(0, 0, 1, 10), (5, 0, 3, 50), (0, 7, 3, 50), (1, 1, 1, 5), (40, 3, 1, 200),
                                            (300, 200, 1, 500), (5000, 3000, 7, 40000), (20000, 9000, 25, 100000),
                                            (3000, 9000, 300, 3000)])
def test_closest_matches_oracle(kw, n1, n2, nb, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 * 7 + n2 * 3 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, 60, neg=(nb == 7))
    g2, s2, e2 = _rand(rng, n2, nb, span, 60, neg=(nb == 7))
    got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, nb, **kw)
    want = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, **kw)
    _check(got, want)


@pytest.mark.parametrize("kw", [dict(k=2), dict(k=2, ignore_upstream=True), dict(k=3, ignore_downstream=True, ignore_overlaps=True)])
def test_closest_strands(kw):
    from bioframe_b200 import engine

    rng = np.random.default_rng(11)
    g1, s1, e1 = _rand(rng, 4000, 5, 30000, 80)
    g2, s2, e2 = _rand(rng, 2500, 5, 30000, 80)
    along = rng.random(4000) < 0.5
    got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, 5, along=along, **kw)
    want = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 5, along=along, **kw)
    _check(got, want)


@pytest.mark.parametrize("k", [1, 3])
@pytest.mark.parametrize("n,nb,span", [(2, 1, 10), (500, 1, 800), (6000, 6, 20000)])
def test_closest_self(k, n, nb, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n + k)
    g, s, e = _rand(rng, n, nb, span, 50)
    got = engine.closest_intidxs(g, s, e, None, None, None, nb, k=k, self_closest=True)
    want = oo.closest_intidxs_coded(g, s, e, None, None, None, nb, k=k, self_closest=True)
    _check(got, want)


def test_closest_heavy_overlap_fanout():
    """one target spanning everything + many short ones: long stabbing back-scans"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(3)
    g2, s2, e2 = _rand(rng, 3000, 2, 50000, 40)
    s2[:4], e2[:4] = 0, 60000
    g1, s1, e1 = _rand(rng, 2000, 2, 50000, 300)
    for kw in (dict(k=1), dict(k=6)):
        got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, 2, **kw)
        want = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 2, **kw)
        _check(got, want)


def test_closest_golden(arrops_golden):
    """golden vectors produced by the reference itself (tie-free targets)"""
    from bioframe_b200 import engine

    from test_oracle_golden import CLOSEST_ALONG, CLOSEST_VARIANTS

    g = arrops_golden
    for name in g["closest_names"]:
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        along = g[f"{name}/along"]
        for tag, kw in CLOSEST_VARIANTS.items():
            a, b, _ = engine.closest_intidxs(None, s1, e1, None, s2, e2, 1, **kw)
            keep = b.cpu().numpy() >= 0  # arrops-level golden has no unmatched rows
            assert np.array_equal(a.cpu().numpy()[keep], g[f"{name}/{tag}_1"]), (name, tag)
            assert np.array_equal(b.cpu().numpy()[keep], g[f"{name}/{tag}_2"]), (name, tag)
        for tag, kw in CLOSEST_ALONG.items():
            a, b, _ = engine.closest_intidxs(None, s1, e1, None, s2, e2, 1, along=along, **kw)
            keep = b.cpu().numpy() >= 0
            assert np.array_equal(a.cpu().numpy()[keep], g[f"{name}/{tag}_1"]), (name, tag)
            assert np.array_equal(b.cpu().numpy()[keep], g[f"{name}/{tag}_2"]), (name, tag)
        for tag, k in (("self_k1", 1), ("self_k3", 3)):
            a, b, _ = engine.closest_intidxs(None, s2, e2, None, None, None, 1, k=k, self_closest=True)
            keep = b.cpu().numpy() >= 0
            assert np.array_equal(a.cpu().numpy()[keep], g[f"{name}/{tag}_1"]), (name, tag)
            assert np.array_equal(b.cpu().numpy()[keep], g[f"{name}/{tag}_2"]), (name, tag)


def test_closest_int64_max_ends():
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    rng = np.random.default_rng(4)
    g1, s1, e1 = _rand(rng, 3000, 3, 10**8, 4000)
    g2, s2, e2 = _rand(rng, 2000, 3, 10**8, 4000)
    e1[::9] = big
    e2[::13] = big
    for kw in (dict(k=2), dict(k=2, ignore_overlaps=True)):
        got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, 3, **kw)
        want = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 3, **kw)
        _check(got, want)


def test_closest_one_giant_target_many_short():
    """a chromosome-long target at the start of the order: the stabbing walk must skip the short
    targets in between block-wise instead of scanning them for every query"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(17)
    n1, n2 = 150_000, 400_000
    g2 = np.zeros(n2, np.int32)
    s2 = rng.integers(1000, 50_000_000, n2).astype(np.int64)
    e2 = s2 + rng.integers(1, 80, n2)
    s2[0], e2[0] = 0, 60_000_000  # overlaps every query
    g1 = np.zeros(n1, np.int32)
    s1 = rng.integers(1000, 50_000_000, n1).astype(np.int64)
    e1 = s1 + rng.integers(1, 200, n1)
    for kw in (dict(k=1), dict(k=3)):
        got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, 1, **kw)
        want = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 1, **kw)
        _check(got, want)
