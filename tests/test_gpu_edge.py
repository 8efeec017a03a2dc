"""Edge cases of the CUDA engine against the oracle: long tie runs (the radix tie path), thousands of
groups (global per-group extent table), empty groups, degenerate sizes, coordinate ranges that do not fit
the 64-bit key (clean error), duplicated inputs."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ops_oracle as oo  # noqa: E402


def _eq(got, want):
    assert got[0].shape[0] == want[0].shape[0]
    assert np.array_equal(got[0].cpu().numpy(), want[0])
    assert np.array_equal(got[1].cpu().numpy(), want[1])


@pytest.mark.parametrize("how", ["inner", "outer"])
def test_long_tie_runs(how):
    """hundreds of rows sharing one start: runs > 32 go through compaction + full-key radix sort"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(1)
    starts = rng.choice(np.array([10, 10, 10, 500, 500, 900, 7000]), 3000).astype(np.int64)
    s1, e1 = starts, starts + rng.integers(0, 300, 3000)
    s2 = rng.choice(np.array([0, 10, 10, 499, 500, 650]), 2000).astype(np.int64)
    e2 = s2 + rng.integers(0, 200, 2000)
    g1 = rng.integers(0, 2, 3000).astype(np.int32)
    g2 = rng.integers(0, 2, 2000).astype(np.int32)
    _eq(engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 2, how=how), oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 2, how))
    res = engine.cluster(g1, s1, e1, 2, min_dist=0)
    ids, cs, ce, cn = oo.cluster_coded(g1, s1, e1, 2, 0)
    assert np.array_equal(res["ids"].cpu().numpy(), ids) and np.array_equal(res["ccount"].cpu().numpy(), cn)


def test_identical_rows():
    from bioframe_b200 import engine

    s = np.full(500, 42, dtype=np.int64)
    e = np.full(500, 50, dtype=np.int64)
    for how in ("inner", "left"):
        _eq(engine.overlap_intidxs(None, s, e, None, s[:70], e[:70], 1, how=how),
            oo.overlap_intidxs_coded(None, s, e, None, s[:70], e[:70], 1, how))
    got = engine.closest_intidxs(None, s, e, None, None, None, 1, k=3, self_closest=True)
    want = oo.closest_intidxs_coded(None, s, e, None, None, None, 1, k=3, self_closest=True)
    _eq(got, want)


@pytest.mark.parametrize("nb", [2049, 5000])
def test_many_groups(nb):
    """more groups than the shared-memory extent table holds; most groups empty on one side"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(nb)
    n1, n2 = 20000, 15000
    g1 = rng.integers(0, nb, n1).astype(np.int32)
    g2 = rng.integers(0, nb // 2, n2).astype(np.int32)  # upper half of the groups has no targets
    s1 = rng.integers(0, 2000, n1).astype(np.int64)
    s2 = rng.integers(0, 2000, n2).astype(np.int64)
    e1, e2 = s1 + rng.integers(0, 400, n1), s2 + rng.integers(0, 400, n2)
    for how in ("inner", "outer"):
        _eq(engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, nb, how=how), oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how))
    _eq(engine.closest_intidxs(g1, s1, e1, g2, s2, e2, nb, k=2), oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, k=2))
    res = engine.cluster(g1, s1, e1, nb, min_dist=3)
    ids, cs, ce, cn = oo.cluster_coded(g1, s1, e1, nb, 3)
    assert np.array_equal(res["ids"].cpu().numpy(), ids) and np.array_equal(res["cstart"].cpu().numpy(), cs)
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, nb).cpu().numpy()
    assert np.array_equal(cov, oo.coverage_coded(g1, s1, e1, g2, s2, e2, nb))


def test_bad_group_code_and_bad_interval():
    from bioframe_b200 import engine

    s = np.array([1, 5], dtype=np.int64)
    e = np.array([4, 9], dtype=np.int64)
    with pytest.raises(ValueError):
        engine.overlap_intidxs(np.array([0, 7], np.int32), s, e, np.array([0, 1], np.int32), s, e, 3)
    with pytest.raises(ValueError):
        engine.overlap_intidxs(None, s, np.array([0, 9], np.int64), None, s, e, 1)  # end < start


def test_span_too_wide_is_a_clean_error():
    """starts spread over (almost) the whole int64 range cannot be packed into a 64-bit key"""
    from bioframe_b200 import engine

    lo, hi = np.iinfo(np.int64).min + 10, np.iinfo(np.int64).max - 10
    s = np.array([lo, 0, hi - 5], dtype=np.int64)
    e = np.array([lo + 3, 8, hi], dtype=np.int64)
    with pytest.raises(OverflowError):
        engine.overlap_intidxs(None, s, e, None, s, e, 1)
    with pytest.raises(OverflowError):
        engine.cluster(None, s, e, 1)


def test_coverage_queries_outside_target_extent():
    from bioframe_b200 import engine

    g2 = np.array([0, 0, 1], np.int32)
    s2 = np.array([100, 150, 10], np.int64)
    e2 = np.array([200, 300, 20], np.int64)
    g1 = np.array([0, 0, 0, 1, 1, 2], np.int32)
    s1 = np.array([-500, 250, 10**12, 0, 15, 5], np.int64)
    e1 = np.array([10**12, 10**9, 10**12 + 5, 12, 10**6, 50], np.int64)
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, 3).cpu().numpy()
    assert np.array_equal(cov, oo.coverage_coded(g1, s1, e1, g2, s2, e2, 3))
    assert list(cov) == [200, 50, 0, 2, 5, 0]


@pytest.mark.parametrize("which", ["both", "first", "second"])
def test_presorted_inputs_skip_the_radix_passes(which):
    """rows already in (group, start) order (typical BED files): the engine sorts nothing and only
    fixes ties; results must not change"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(33)

    def make(n, sort):
        g = rng.integers(0, 6, n).astype(np.int32)
        s = rng.integers(0, 3000, n).astype(np.int64)  # dense: plenty of equal starts with unsorted ends
        e = s + rng.integers(0, 80, n)
        if sort:
            o = np.lexsort([s, g])
            g, s, e = g[o], s[o], e[o]
        return g, s, e

    g1, s1, e1 = make(30000, which in ("both", "first"))
    g2, s2, e2 = make(20000, which in ("both", "second"))
    for how in ("inner", "outer"):
        _eq(engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 6, how=how), oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 6, how))
    _eq(engine.closest_intidxs(g1, s1, e1, g2, s2, e2, 6, k=3), oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 6, k=3))
    cnt = engine.count_overlaps(g1, s1, e1, g2, s2, e2, 6).cpu().numpy()
    i, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 6, how="inner")
    assert np.array_equal(cnt, np.bincount(i, minlength=len(s1)))
    for md in (0, None):
        res = engine.cluster(g1, s1, e1, 6, min_dist=md)
        ids, cs, ce, cn = oo.cluster_coded(g1, s1, e1, 6, md)
        assert np.array_equal(res["ids"].cpu().numpy(), ids) and np.array_equal(res["cend"].cpu().numpy(), ce)
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, 6).cpu().numpy()
    assert np.array_equal(cov, oo.coverage_coded(g1, s1, e1, g2, s2, e2, 6))


def test_many_groups_times_long_intervals_take_group_batches():
    """~1000 (chrom, on) groups of hg38-sized spans need ~38 position bits, one chromosome-long interval
    28 length bits more: no single 64-bit key layout holds all groups, so the engine walks over group batches --
    same rows as the oracle, for every op."""
    from bioframe_b200 import engine
    from oracle import ops_oracle as oo

    rng = np.random.default_rng(5)
    nb, span = 1000, 250_000_000
    n1, n2 = 6000, 5000
    g1, g2 = rng.integers(0, nb, n1).astype(np.int32), rng.integers(0, nb, n2).astype(np.int32)
    s1, s2 = rng.integers(0, span, n1), rng.integers(0, span, n2)
    e1 = s1 + rng.integers(0, 3_000_000, n1)
    e2 = s2 + rng.integers(0, 3_000_000, n2)
    e1[::50] = s1[::50] + span  # chromosome-long intervals: 28 length bits
    e2[::70] = s2[::70] + span
    d = [torch.from_numpy(x).cuda() for x in (g1, s1, e1, g2, s2, e2)]
    for how in ("inner", "outer"):
        a, b, _ = engine.overlap_intidxs(*d, nb, how=how)
        w1, w2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how=how)
        assert np.array_equal(a.cpu().numpy(), w1) and np.array_equal(b.cpu().numpy(), w2)
    cnt = engine.count_overlaps(*d, nb).cpu().numpy()
    assert np.array_equal(cnt, np.bincount(w1[(w1 >= 0) & (w2 >= 0)], minlength=n1))
    assert np.array_equal(engine.coverage(*d, nb).cpu().numpy(), oo.coverage_coded(g1, s1, e1, g2, s2, e2, nb))
    res = engine.cluster(d[0], d[1], d[2], nb, min_dist=0)
    ids, cs, ce, cn = oo.cluster_coded(g1, s1, e1, nb, 0)
    assert np.array_equal(res["ids"].cpu().numpy(), ids) and np.array_equal(res["cstart"].cpu().numpy(), cs)
    assert np.array_equal(res["cend"].cpu().numpy(), ce) and np.array_equal(res["ccount"].cpu().numpy(), cn)
    row, ps, pe, sub = engine.subtract(*d, nb)
    wr, wps, wpe, wsub = oo.subtract_coded(g1, s1, e1, g2, s2, e2, nb)
    assert np.array_equal(row.cpu().numpy(), wr) and np.array_equal(ps.cpu().numpy(), wps)
    assert np.array_equal(pe.cpu().numpy(), wpe) and np.array_equal(sub.cpu().numpy(), wsub)


def test_coverage_of_complement_output():
    """`coverage(df1, complement(df2))` -- the complement ends every chromosome at INT64_MAX
    (ops.py:1604), which the coverage key layout cannot hold as it is."""
    import pandas as pd

    import bioframe_b200 as bf
    from oracle import ops_oracle as oo

    rng = np.random.default_rng(9)
    chroms = np.array(["chr1", "chr2", "chr3"], dtype=object)

    def frame(n):
        s = rng.integers(0, 100_000, n)
        return pd.DataFrame({"chrom": rng.choice(chroms, n), "start": s, "end": s + rng.integers(1, 3000, n)})

    df1, df2 = frame(4000), frame(500)
    gaps = bf.complement(df2)
    assert int(gaps["end"].max()) == np.iinfo(np.int64).max
    got = bf.coverage(df1, gaps)
    code = {c: i for i, c in enumerate(chroms)}
    g1 = df1["chrom"].map(code).to_numpy().astype(np.int32)
    g2 = gaps["chrom"].map(code).to_numpy().astype(np.int32)
    want = oo.coverage_coded(g1, df1["start"].to_numpy(), df1["end"].to_numpy(), g2, gaps["start"].to_numpy(),
                             gaps["end"].to_numpy(), 3)
    assert np.array_equal(got["coverage"].to_numpy(), want)
    # and the two coverages of a query add up to its length where the view covers it
    both = bf.coverage(df1, df2)["coverage"].to_numpy() + got["coverage"].to_numpy()
    assert np.array_equal(both, (df1["end"] - df1["start"]).to_numpy())
