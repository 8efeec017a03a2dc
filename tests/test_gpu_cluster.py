"""CUDA cluster / merge / coverage / complement against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import arrops_oracle as ao  # noqa: E402
from oracle import ops_oracle as oo  # noqa: E402

BIG = np.iinfo(np.int64).max


def _rand(rng, n, nb, span, maxlen, p_point=0.1, neg=False):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(-span if neg else 0, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return g, s, s + ln


CASES = [(1, 1, 10), (2, 1, 10), (300, 1, 500), (5000, 7, 4000), (20000, 25, 100000), (30000, 300, 3000),
         (70000, 3, 2000000)]


@pytest.mark.parametrize("min_dist", [0, None, 1, 7, 1000])
@pytest.mark.parametrize("n,nb,span", CASES)
def test_cluster_matches_oracle(n, nb, span, min_dist):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n + nb)
    g, s, e = _rand(rng, n, nb, span, 60, neg=(nb == 7))
    res = engine.cluster(g, s, e, nb, min_dist=min_dist, want_row_spans=True)
    ids, cs, ce, cn = oo.cluster_coded(g, s, e, nb, min_dist)
    assert res["n_clusters"] == len(cs)
    assert np.array_equal(res["ids"].cpu().numpy(), ids)
    assert np.array_equal(res["cstart"].cpu().numpy(), cs)
    assert np.array_equal(res["cend"].cpu().numpy(), ce)
    assert np.array_equal(res["ccount"].cpu().numpy(), cn)
    assert np.array_equal(res["cgrp"].cpu().numpy(), oo.cluster_coded.last_groups)
    assert np.array_equal(res["row_start"].cpu().numpy(), cs[ids])
    assert np.array_equal(res["row_end"].cpu().numpy(), ce[ids])


def test_cluster_empty_and_table_only():
    from bioframe_b200 import engine

    z = np.empty(0, np.int64)
    res = engine.cluster(np.empty(0, np.int32), z, z, 3)
    assert res["n_clusters"] == 0 and res["ids"].numel() == 0
    rng = np.random.default_rng(5)
    g, s, e = _rand(rng, 4000, 5, 3000, 40)
    res = engine.cluster(g, s, e, 5, want_ids=False)
    _, cs, ce, cn = oo.cluster_coded(g, s, e, 5, 0)
    assert res["ids"] is None
    assert np.array_equal(res["cstart"].cpu().numpy(), cs) and np.array_equal(res["ccount"].cpu().numpy(), cn)


def test_merge_golden(arrops_golden):
    from bioframe_b200 import engine

    g = arrops_golden
    for name in g["names"]:
        if name == "wide":
            continue
        s1, e1 = g[f"{name}/s1"], g[f"{name}/e1"]
        if len(s1) == 0:
            continue
        for tag, md in (("0", 0), ("none", None), ("2", 2), ("75", 75)):
            res = engine.cluster(None, s1, e1, 1, min_dist=md)
            assert np.array_equal(res["ids"].cpu().numpy(), g[f"{name}/merge_{tag}_ids"])
            assert np.array_equal(res["cstart"].cpu().numpy(), g[f"{name}/merge_{tag}_s"])
            assert np.array_equal(res["cend"].cpu().numpy(), g[f"{name}/merge_{tag}_e"])


@pytest.mark.parametrize("n1,n2,nb,span", [(0, 5, 2, 50), (5, 0, 2, 50), (300, 200, 1, 500), (5000, 3000, 7, 4000),
                                            (20000, 9000, 25, 100000), (9000, 30000, 300, 3000)])
def test_coverage_matches_oracle(n1, n2, nb, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 + 3 * n2 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, 200, neg=(nb == 7))
    g2, s2, e2 = _rand(rng, n2, nb, span, 60, neg=(nb == 7))
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, nb).cpu().numpy()
    assert np.array_equal(cov, oo.coverage_coded(g1, s1, e1, g2, s2, e2, nb))


@pytest.mark.parametrize("n,nr,span", [(0, 3, 100), (1, 1, 10), (400, 1, 500), (6000, 5, 4000), (20000, 40, 3000)])
def test_complement_matches_oracle(n, nr, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n + nr)
    g, s, e = _rand(rng, n, nr, span, 30, p_point=0.0)
    bounds = []
    for r in range(nr):
        lo = int(rng.integers(-10, span // 3))
        hi = int(rng.integers(span // 2, span + 50)) if r % 3 else BIG
        bounds.append((lo, hi))
    # the reference clips intervals to their view region before complementing (ops.py:1614-1622)
    b0 = np.array([b[0] for b in bounds], dtype=np.int64)
    b1 = np.array([b[1] for b in bounds], dtype=np.int64)
    if n:
        keep = (e > b0[g]) & (s < b1[g])
        g, s, e = g[keep], np.maximum(s[keep], b0[g[keep]]), np.minimum(e[keep], b1[g[keep]])
    reg, gs, ge = engine.complement(g, s, e, b0, b1)
    oc, os_, oe = oo.complement_core(g, s, e, bounds)
    assert np.array_equal(reg.cpu().numpy(), oc)
    assert np.array_equal(gs.cpu().numpy(), os_)
    assert np.array_equal(ge.cpu().numpy(), oe)


def test_complement_golden(arrops_golden):
    from bioframe_b200 import engine

    g = arrops_golden
    for name in g["names"]:
        if name == "wide":
            continue
        s1, e1 = g[f"{name}/s1"], g[f"{name}/e1"]
        if len(s1) == 0:
            continue
        for tag in ("def", "b", "out"):
            b = [int(x) for x in g[f"{name}/compl_{tag}_bounds"]]
            _, gs, ge = engine.complement(None, s1, e1, np.array([b[0]]), np.array([b[1]]))
            assert np.array_equal(gs.cpu().numpy(), g[f"{name}/compl_{tag}_s"]), (name, tag)
            assert np.array_equal(ge.cpu().numpy(), g[f"{name}/compl_{tag}_e"]), (name, tag)


@pytest.mark.parametrize("closed", [False, True])
@pytest.mark.parametrize("n1,n2,nb,span", [(0, 5, 2, 50), (5, 0, 2, 50), (300, 200, 1, 500), (5000, 3000, 7, 4000),
                                            (20000, 9000, 25, 100000), (9000, 30000, 300, 3000)])
def test_count_overlaps_matches_oracle(n1, n2, nb, span, closed):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 + 5 * n2 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, 200, neg=(nb == 7))
    g2, s2, e2 = _rand(rng, n2, nb, span, 60, neg=(nb == 7))
    got = engine.count_overlaps(g1, s1, e1, g2, s2, e2, nb, closed=closed).cpu().numpy()
    i, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how="inner", closed=closed)
    assert np.array_equal(got, np.bincount(i, minlength=n1))


def test_count_overlaps_int64_max():
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    rng = np.random.default_rng(8)
    g1, s1, e1 = _rand(rng, 3000, 3, 10**8, 4000)
    g2, s2, e2 = _rand(rng, 2000, 3, 10**8, 4000)
    e1[::9] = big
    e2[::13] = big
    got = engine.count_overlaps(g1, s1, e1, g2, s2, e2, 3).cpu().numpy()
    i, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 3, how="inner")
    assert np.array_equal(got, np.bincount(i, minlength=3000))


@pytest.mark.parametrize("min_dist", [0, None, 30])
def test_cluster_merge_int64_max_ends(min_dist):
    """the last gap of a complement ends at INT64_MAX: cluster / merge must take such ends.  With
    min_dist > 0 the reference computes `running_max + min_dist` in int64, which wraps for such an end and
    opens a spurious cluster after it; here the end simply reaches past everything, i.e. the
    answer the reference gives when the huge ends are replaced by a large value that does not overflow."""
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    rng = np.random.default_rng(12)
    g, s, e = _rand(rng, 6000, 4, 10**6, 300)
    huge = np.zeros(len(e), dtype=bool)
    huge[::37] = True
    e[huge] = big
    e[5] = big - 9
    huge[5] = True
    res = engine.cluster(g, s, e, 4, min_dist=min_dist, want_row_spans=True)
    safe = np.where(huge, 2**61, e)
    ids, cs, ce, cn = oo.cluster_coded(g, s, safe, 4, min_dist)
    if not min_dist:  # no overflow in the reference either: identical including the ends
        ids0, cs0, ce0, cn0 = oo.cluster_coded(g, s, e, 4, min_dist)
        assert np.array_equal(ids, ids0) and np.array_equal(cn, cn0)
    true_end = np.full(len(cs), np.iinfo(np.int64).min)
    np.maximum.at(true_end, ids, e)
    assert np.array_equal(res["ids"].cpu().numpy(), ids)
    assert np.array_equal(res["cstart"].cpu().numpy(), cs)
    assert np.array_equal(res["cend"].cpu().numpy(), true_end)
    assert np.array_equal(res["ccount"].cpu().numpy(), cn)
    assert np.array_equal(res["row_end"].cpu().numpy(), true_end[ids])


def test_complement_with_int64_max_interval():
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    s = np.array([10, 50, 200], dtype=np.int64)
    e = np.array([20, big, 300], dtype=np.int64)
    _, gs, ge = engine.complement(None, s, e, np.array([0]), np.array([big]))
    cs, ce = ao.complement_intervals(s, e, bounds=(0, big))
    assert np.array_equal(gs.cpu().numpy(), cs) and np.array_equal(ge.cpu().numpy(), ce)


def test_coverage_with_int64_max_ends():
    """targets that end at INT64_MAX (what complement / subtract / trim put at the end of a chromosome,
    ops.py:1309, 1499, 1604) do not fit the coverage key layout as they are: the engine clips every coordinate to the
    window in which both sets have intervals and answers like the reference"""
    from bioframe_b200 import engine
    from oracle import ops_oracle as oo

    big = np.iinfo(np.int64).max
    s1, e1 = np.array([5, 0, 40, 7]), np.array([9, 3, 45, 7])
    s2, e2 = np.array([1, 4, 43]), np.array([big, 6, big])
    got = engine.coverage(None, s1, e1, None, s2, e2, 1).cpu().numpy()
    assert np.array_equal(got, oo.coverage_coded(None, s1, e1, None, s2, e2, 1))
    g1, g2 = np.array([0, 1, 1, 0], dtype=np.int32), np.array([1, 0, 1], dtype=np.int32)
    got = engine.coverage(g1, s1, e1, g2, s2, e2, 2).cpu().numpy()
    assert np.array_equal(got, oo.coverage_coded(g1, s1, e1, g2, s2, e2, 2))
