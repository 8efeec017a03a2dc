"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run
1e8 x 1e7 in seconds): every emitted pair really overlaps, the pairs per query equal the counts of the
independent count_overlaps kernel (two binary searches per query, no shared code with the emit path),
unpaired queries are exactly the zero-count ones, the keep_order sort is sorted and a permutation, cluster /
merge are consistent and idempotent, coverage obeys its bounds and equals the interval length on the merged
set itself, closest(k=1) agrees with a brute-force scan on sampled queries.  All checks run on the device
with torch ops; nothing here imports the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N1, N2, NB = 100_000_000, 10_000_000, 24


@pytest.fixture(scope="module")
def sets():
    import bench

    (q, t) = bench.workload(N1, N2, 0)
    dev = torch.device("cuda", 0)
    return [torch.from_numpy(a).to(dev) for a in (*q, *t)]


def _adj(s, e):
    return e + (e == s).to(e.dtype)


def test_overlap_left_full_size_properties(sets):
    from bioframe_b200 import engine

    g1, s1, e1, g2, s2, e2 = sets
    ev1, ev2, n_pairs = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, NB, how="left")
    paired = ev2 >= 0
    assert int(paired.sum().item()) == n_pairs
    assert bool((ev1 >= 0).all().item())
    # 1. every pair overlaps, in the same group (checked in slices to bound temporary memory)
    step = 1 << 27
    for lo in range(0, ev1.numel(), step):
        a, b, m = ev1[lo:lo + step], ev2[lo:lo + step], paired[lo:lo + step]
        a, b = a[m], b[m]
        assert bool((g1[a] == g2[b]).all().item())
        assert bool((s1[a] < _adj(s2[b], e2[b])).all().item())
        assert bool((s2[b] < _adj(s1[a], e1[a])).all().item())
    # 2. pairs per query == independent count kernel; 3. unpaired rows == zero-count queries, once each
    per_query = torch.bincount(ev1[paired], minlength=N1)
    counts = engine.count_overlaps(g1, s1, e1, g2, s2, e2, NB)
    assert torch.equal(per_query, counts)
    lone = torch.bincount(ev1[~paired], minlength=N1)
    assert torch.equal(lone, (counts == 0).to(lone.dtype))
    assert ev1.numel() == int(counts.sum().item()) + int((counts == 0).sum().item())
    del per_query, lone, counts
    # 4. keep_order: sorted by (row1, row2) and a permutation (order-independent checksums unchanged)
    chk = (int(ev1.sum().item()), int((ev2 * 3 + ev1 % 7).sum().item()))
    engine.sort_pairs(ev1, ev2, N1 - 1, N2 - 1)
    assert chk == (int(ev1.sum().item()), int((ev2 * 3 + ev1 % 7).sum().item()))
    assert bool(((ev1[1:] > ev1[:-1]) | ((ev1[1:] == ev1[:-1]) & (ev2[1:] > ev2[:-1]))).all().item())


def test_overlap_inner_pair_set_is_symmetric(sets):
    """swapping the two sets yields the same pair set (A/B blocks trade places)"""
    from bioframe_b200 import engine

    g1, s1, e1, g2, s2, e2 = sets
    n = 20_000_000  # a fifth of the queries keeps the two pair lists + sort scratch well inside HBM
    a1, a2, na = engine.overlap_intidxs(g1[:n], s1[:n], e1[:n], g2, s2, e2, NB, how="inner")
    b2, b1, nb = engine.overlap_intidxs(g2, s2, e2, g1[:n], s1[:n], e1[:n], NB, how="inner")
    assert na == nb == a1.numel() == b1.numel()
    engine.sort_pairs(a1, a2, n - 1, N2 - 1)
    engine.sort_pairs(b1, b2, n - 1, N2 - 1)
    assert torch.equal(a1, b1) and torch.equal(a2, b2)


def test_cluster_merge_full_size_properties():
    from bioframe_b200 import engine, synth

    g, s, e = synth.make_intervals(N1, 1, "geom1k")  # sparse enough for millions of clusters
    dev = torch.device("cuda", 0)
    g, s, e = (torch.from_numpy(x).to(dev) for x in (g, s, e))
    res = engine.cluster(g, s, e, NB, min_dist=0, want_row_spans=True)
    ids, C = res["ids"], res["n_clusters"]
    assert int(res["ccount"].sum().item()) == N1
    assert torch.equal(torch.bincount(ids, minlength=C), res["ccount"])
    assert bool((res["row_start"] <= s).all().item()) and bool((e <= res["row_end"]).all().item())
    assert bool((res["cgrp"].long()[ids] == g.long()).all().item())
    # clusters of one group are disjoint, non-adjacent (min_dist=0 merges touching intervals) and ascending
    same = res["cgrp"][1:] == res["cgrp"][:-1]
    assert bool((res["cstart"][1:][same] > res["cend"][:-1][same]).all().item())
    assert bool((res["cgrp"][1:] >= res["cgrp"][:-1]).all().item())
    # idempotence: merging the merged intervals changes nothing
    again = engine.cluster(res["cgrp"], res["cstart"], res["cend"], NB, min_dist=0)
    assert again["n_clusters"] == C
    assert torch.equal(again["cstart"], res["cstart"]) and torch.equal(again["cend"], res["cend"])
    assert torch.equal(again["ids"], torch.arange(C, device=dev))
    # coverage of the merged set by the original set == its own length; coverage <= length in general
    cov = engine.coverage(res["cgrp"], res["cstart"], res["cend"], g, s, e, NB)
    assert torch.equal(cov, res["cend"] - res["cstart"])


def test_coverage_and_counts_bounds(sets):
    from bioframe_b200 import engine

    g1, s1, e1, g2, s2, e2 = sets
    cov = engine.coverage(g1, s1, e1, g2, s2, e2, NB)
    assert bool((cov >= 0).all().item()) and bool((cov <= e1 - s1).all().item())
    counts = engine.count_overlaps(g1, s1, e1, g2, s2, e2, NB)
    # a query with covered bases overlaps something; points aside, a query that overlaps nothing has coverage 0
    assert bool((counts[cov > 0] > 0).all().item())
    assert bool((cov[counts == 0] == 0).all().item())
    # total coverage two ways: sum over queries vs. by symmetry nothing to compare against -- instead check
    # additivity on split queries: cov([s, e)) == cov([s, m)) + cov([m, e))
    m = s1 + (e1 - s1) // 2
    left = engine.coverage(g1, s1, m, g2, s2, e2, NB)
    right = engine.coverage(g1, m, e1, g2, s2, e2, NB)
    assert torch.equal(cov, left + right)


def test_closest_k1_brute_force_sample():
    from bioframe_b200 import engine, synth

    n1, n2 = 10_000_000, 1_000_000
    g1, s1, e1 = synth.make_intervals(n1, 1, "geom1k")
    g2, s2, e2 = synth.make_intervals(n2, 2, "geom1k")
    ev1, ev2, n_matched = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, NB, k=1)
    ev1, ev2 = ev1.cpu().numpy(), ev2.cpu().numpy()
    assert n_matched == n1 and len(ev1) == n1  # every chromosome has targets: one row per query
    assert np.array_equal(np.sort(ev1), np.arange(n1))
    assert np.array_equal(g1[ev1], g2[ev2])
    # reported distance == brute-force minimum distance over the query's chromosome, on a sample
    dist = np.maximum(0, np.maximum(s1[ev1] - e2[ev2], s2[ev2] - e1[ev1]))
    rng = np.random.default_rng(0)
    by_chrom = [np.flatnonzero(g2 == c) for c in range(NB)]
    for r in rng.integers(0, n1, 300):
        q = ev1[r]
        t = by_chrom[g1[q]]
        best = np.maximum(0, np.maximum(s1[q] - e2[t], s2[t] - e1[q])).min()
        assert dist[r] == best


def test_ordered_left_overlap_equals_sorted_pair_list_at_full_size(sets):
    """Two independent routes to the keep_order result at 1e8 x 1e7: pairs born in row order
    (bf_overlap_ordered_*: per-query probes + running-max walk) against all pairs + radix sort of the pair
    list (bf_overlap_* + bf_sort_pairs).  They share the sort of the targets and nothing else."""
    from bioframe_b200 import engine

    g1, s1, e1, g2, s2, e2 = sets
    a, b, biggest = engine.overlap_ordered(g1, s1, e1, g2, s2, e2, NB)
    ev1, ev2, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, NB, how="left")
    ev1, ev2 = engine.sort_pairs(ev1, ev2, N1 - 1, N2 - 1)
    assert a.numel() == ev1.numel()
    assert torch.equal(a, ev1) and torch.equal(b, ev2)
    assert 0 < biggest < 100_000
