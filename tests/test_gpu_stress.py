"""Stress of the parts whose correctness rests on inter-block protocols (compute-sanitizer's racecheck is closed
on this pool): the decoupled look-back of the radix pass (relaxed atomics on per-(tile, digit) status words, tiles
taken by ticket), the in-place finishing pass (a run is owned by the tile of its first row) and the ticketed
running-max scans -- many repetitions over sizes around every tile boundary, each compared with an independent
torch sort / the oracle.  A race shows up as a mismatch in some repetition, not as a crash."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _sorted_reference(g, s, e):
    """(group, start, end', row) order = np.lexsort([ends', starts]) per group (arrops.py:332-335)"""
    ee = e + (e == s)
    return np.lexsort([np.arange(len(s)), ee, s, g])


@pytest.mark.parametrize("n", [1, 255, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 65_537, 300_001, 1_048_577, 5_000_011])
def test_sort_order_repeated(n):
    """the permutation of `prepare_set` against a host lexsort, 12 repetitions per size with fresh data (different
    tile timing every time); coordinates dense enough for many equal starts and long tie runs"""
    from bioframe_b200 import engine

    rng = np.random.default_rng(n)
    reps = 12 if n < 1_000_000 else 4
    for rep in range(reps):
        span = int(rng.choice([50, 5_000, 3_000_000]))
        g = rng.integers(0, 7, n).astype(np.int32)
        s = rng.integers(0, span, n).astype(np.int64)
        e = s + rng.integers(0, 40, n)
        d = [torch.from_numpy(x).cuda() for x in (g, s, e)]
        words, pres = engine.measure_extent(*d, 7)
        st = engine.prepare_set(words, *d, 7, presorted=pres)
        got = st.ids.cpu().numpy().astype(np.int64)
        want = _sorted_reference(g, s, e)
        assert np.array_equal(got, want), (n, rep, span)


def test_overlap_repeated_same_input_is_deterministic():
    """the same call 30 times: every result identical to the first (look-back / ticket order must not leak into
    the output) and equal to the oracle"""
    from bioframe_b200 import engine, synth
    from oracle import ops_oracle as oo

    q = synth.make_intervals(400_000, 1, "geom2k")
    t = synth.make_peak_targets(60_000, 2, n_hotspots=1500)
    s = lambda x: x // 100  # noqa: E731  (denser: more pairs, more ties)
    q = (q[0], s(q[1]), s(q[2]) + 1)
    t = (t[0], s(t[1]), s(t[2]) + 1)
    d = [torch.from_numpy(x).cuda() for x in (*q, *t)]
    w1, w2 = oo.overlap_intidxs_coded(*q, *t, 24, how="outer")
    for rep in range(30):
        a, b, _ = engine.overlap_intidxs(*d, 24, how="outer")
        assert np.array_equal(a.cpu().numpy(), w1) and np.array_equal(b.cpu().numpy(), w2), rep


def test_cluster_and_closest_repeated():
    from bioframe_b200 import engine, synth
    from oracle import ops_oracle as oo

    g, s, e = synth.make_intervals(700_000, 3, "pareto")
    d = [torch.from_numpy(x).cuda() for x in (g, s, e)]
    ids, cs, ce, cn = oo.cluster_coded(g, s, e, 24, 0)
    tq = synth.make_intervals(300_000, 4, "geom1k")
    tt = synth.make_tie_free(50_000, 5, "geom1k")
    dq = [torch.from_numpy(x).cuda() for x in (*tq, *tt)]
    c1, c2 = oo.closest_intidxs_coded(*tq, *tt, 24, k=2)
    for rep in range(15):
        res = engine.cluster(*d, 24, min_dist=0)
        assert np.array_equal(res["ids"].cpu().numpy(), ids) and np.array_equal(res["ccount"].cpu().numpy(), cn), rep
        a, b, _ = engine.closest_intidxs(*dq, 24, k=2)
        assert np.array_equal(a.cpu().numpy(), c1) and np.array_equal(b.cpu().numpy(), c2), rep
