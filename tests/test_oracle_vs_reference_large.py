"""Pins the numpy oracle port against the REAL reference at sizes the golden fixtures cannot hold
(SURVEY.md §8c "parity unpinned (v): anything at >= 1e4 rows"): cfg0-sized overlap (1e6 x 1e6, all four joins),
closest at 2e6 x 1e6 on tie-free targets, cluster / merge at 2e6 heavy-tailed rows, coverage -- array-equal in
the reference's row order.  CPU only; skipped where neither /root/reference nor the vendored copy exists."""
import numpy as np
import pytest

import refcheck
from bioframe_b200 import synth
from oracle import ops_oracle as oo

pytestmark = pytest.mark.skipif(not refcheck.have_reference(), reason="reference not available")
NB = 24


def test_overlap_cfg0_size():
    q = synth.make_intervals(1_000_000, 1, "geom1k")
    t = synth.make_intervals(1_000_000, 2, "geom1k")
    for how in ("inner", "left", "right", "outer"):
        a, b, kind = refcheck.overlap(*q, *t, NB, how)
        assert kind == "reference"
        c, d = oo.overlap_intidxs_coded(*q, *t, NB, how=how)
        assert np.array_equal(a, c) and np.array_equal(b, d), how


def test_closest_tie_free():
    q = synth.make_intervals(2_000_000, 1, "geom1k")
    t = synth.make_tie_free(1_000_000, 2, "geom1k")
    for kw in (dict(k=1), dict(k=2, ignore_overlaps=True)):
        a, b, _ = refcheck.closest(*q, *t, NB, **kw)
        c, d = oo.closest_intidxs_coded(*q, *t, NB, **kw)
        assert np.array_equal(a, c) and np.array_equal(b, d), kw


def test_cluster_merge_heavy_tail():
    g, s, e = synth.make_intervals(2_000_000, 1, "pareto")
    for md in (0, None, 1000):
        ids, rs, re_, (tg, ts, te, tn), _ = refcheck.cluster(g, s, e, NB, md)
        oi, cs, ce, cn = oo.cluster_coded(g, s, e, NB, md)
        assert np.array_equal(ids, oi) and np.array_equal(rs, cs[oi]) and np.array_equal(re_, ce[oi])
        assert np.array_equal(ts, cs) and np.array_equal(te, ce) and np.array_equal(tn, cn)
        assert np.array_equal(tg, oo.cluster_coded.last_groups)


def test_coverage_peaks():
    q = synth.make_intervals(1_000_000, 1, "geom2k")
    t = synth.make_peak_targets(100_000, 2, n_hotspots=2000)
    cov, _ = refcheck.coverage(*q, *t, NB)
    assert np.array_equal(cov, oo.coverage_coded(*q, *t, NB))
