"""CUDA subtract (bf_subtract_*) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ops_oracle as oo  # noqa: E402


def _rand(rng, n, nb, span, maxlen, p_point=0.1, lo=0):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(lo, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return g, s, s + ln


def _check(g1, s1, e1, g2, s2, e2, nb):
    from bioframe_b200 import engine

    row, ps, pe, sub = engine.subtract(g1, s1, e1, g2, s2, e2, nb)
    want = oo.subtract_coded(g1, s1, e1, g2, s2, e2, nb)
    for got, ref, what in zip((row, ps, pe, sub), want, ("row", "start", "end", "sub_index")):
        assert np.array_equal(got.cpu().numpy(), ref), what
    return len(want[0])


@pytest.mark.parametrize("n1,n2,nb,span,maxlen", [
    (0, 5, 2, 50, 10), (5, 0, 2, 50, 10), (1, 1, 1, 10, 5), (300, 200, 1, 500, 30), (5000, 3000, 7, 4000, 60),
    (20000, 9000, 25, 100000, 400), (9000, 30000, 300, 3000, 20), (40000, 500, 3, 2000000, 90000),
    (3000, 40000, 2, 50000, 8),
])
def test_subtract_matches_oracle(n1, n2, nb, span, maxlen):
    rng = np.random.default_rng(n1 + 3 * n2 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, maxlen)
    g2, s2, e2 = _rand(rng, n2, nb, span, maxlen)
    _check(g1, s1, e1, g2, s2, e2, nb)


def test_subtract_negative_coordinates_and_zero_starts():
    # the view starts at 0: negative targets are clipped or dropped, queries keep only what lies in a gap
    rng = np.random.default_rng(11)
    g1, s1, e1 = _rand(rng, 6000, 4, 3000, 50, lo=-600)
    g2, s2, e2 = _rand(rng, 2500, 4, 3000, 40, lo=-900)
    s2[:40] = 0  # merged intervals starting exactly at 0: no leading gap
    e2[:40] = np.maximum(e2[:40], 0)
    e2[40:60] = s2[40:60] = 0  # points at 0
    assert _check(g1, s1, e1, g2, s2, e2, 4) > 0


def test_subtract_groups_without_targets_and_bad_codes():
    from bioframe_b200 import engine

    rng = np.random.default_rng(3)
    g1, s1, e1 = _rand(rng, 4000, 6, 2000, 30)
    g2, s2, e2 = _rand(rng, 1500, 3, 2000, 30)  # groups 3..5 have no targets: rows come back whole
    _check(g1, s1, e1, g2, s2, e2, 6)
    g1[::5] = -1  # rows without a group (NA chromosome) produce nothing
    row, *_ = engine.subtract(g1, s1, e1, g2, s2, e2, 6)
    assert not np.isin(row.cpu().numpy(), np.flatnonzero(g1 < 0)).any()


def test_subtract_one_query_many_pieces():
    s2 = np.arange(10, 400010, 4, dtype=np.int64)
    e2 = s2 + 2
    s1 = np.array([0, 11, 399990, 5], dtype=np.int64)
    e1 = np.array([500000, 12, 400100, 5], dtype=np.int64)
    n = _check(None, s1, e1, None, s2, e2, 1)
    assert n > 100000


def test_subtract_pieces_add_up_with_coverage():
    """Size-independent property at 2e6 x 3e5: piece lengths of a row sum to its length minus its coverage,
    pieces are ordered and none of them overlaps a target."""
    from bioframe_b200 import engine

    rng = np.random.default_rng(21)
    n1, n2, nb = 2_000_000, 300_000, 24
    g1, s1, e1 = _rand(rng, n1, nb, 50_000_000, 4000, p_point=0.0)
    g2, s2, e2 = _rand(rng, n2, nb, 50_000_000, 60000, p_point=0.0)
    row, ps, pe, sub = engine.subtract(g1, s1, e1, g2, s2, e2, nb)
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, nb)
    t = lambda a: torch.as_tensor(a, device=row.device)  # noqa: E731
    total = torch.zeros(n1, dtype=torch.int64, device=row.device).index_add_(0, row, pe - ps)
    assert torch.equal(total, t(e1) - t(s1) - cov)
    assert bool((pe > ps).all())
    assert bool((row[1:] >= row[:-1]).all())
    same = row[1:] == row[:-1]
    assert bool((ps[1:][same] > pe[:-1][same]).all())
    assert bool((sub[1:][same] == sub[:-1][same] + 1).all()) and bool((sub[1:][~same] == 0).all())
    counts = engine.count_overlaps(g1[row.cpu().numpy()], ps, pe, g2, s2, e2, nb)
    assert int(counts.sum()) == 0
