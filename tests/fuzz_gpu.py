#!/usr/bin/env python
"""Randomised differential run of every engine entry point against the oracle (not collected by pytest:
run by hand on the GPU box, `python tests/fuzz_gpu.py [seconds] [seed]`).  Varies sizes, group counts,
coordinate spans (dense ties ... sparse), interval lengths, point fractions, negative coordinates,
INT64_MAX ends, join kinds, closed, k / ignore flags / strands."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from bioframe_b200 import engine  # noqa: E402
from oracle import ops_oracle as oo  # noqa: E402

BIG = np.iinfo(np.int64).max


def make(rng, n, nb, span, maxlen, p_point, neg, p_big):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(-span if neg else 0, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    e = s + ln
    if p_big:
        e[rng.random(n) < p_big] = BIG
    return g, s, e


def same(got, want, what, cfg):
    ok = all(np.array_equal(np.asarray(a.cpu() if hasattr(a, "cpu") else a), b) for a, b in zip(got, want))
    if not ok:
        print("MISMATCH", what, cfg, flush=True)
        raise SystemExit(1)


def run(budget=120.0, seed=0, max_rounds=None, sizes1=(0, 1, 7, 300, 2500, 9000, 40000),
        sizes2=(0, 1, 5, 200, 3000, 10000, 30000)):
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    it = 0
    while time.time() < t_end and (max_rounds is None or it < max_rounds):
        it += 1
        nb = int(rng.choice([1, 1, 2, 5, 24, 100, 3000]))
        n1 = int(rng.choice(sizes1))
        n2 = int(rng.choice(sizes2))
        # groups are laid end to end on one axis: span x groups (+ length bits) must fit the 64-bit key
        span = int(rng.choice([5, 60, 1000, 50000, 10**7, 10**12] if nb <= 24 else [5, 60, 1000, 50000, 10**7]))
        maxlen = int(rng.choice([2, 10, 100, 5000]))
        p_point = float(rng.choice([0.0, 0.1, 0.6]))
        neg = bool(rng.random() < 0.3)
        # INT64_MAX ends are packed by clamping to (largest start + 1): that needs span + length bits <= 64
        # (DESIGN.md 2), fine for genome-sized spans, not for spans of 1e12 spread over many groups
        p_big = float(rng.choice([0.0, 0.0, 0.02])) if span <= 10**7 else 0.0
        cfg = dict(it=it, nb=nb, n1=n1, n2=n2, span=span, maxlen=maxlen, p_point=p_point, neg=neg, p_big=p_big)
        g1, s1, e1 = make(rng, n1, nb, span, maxlen, p_point, neg, p_big)
        g2, s2, e2 = make(rng, n2, nb, span, maxlen, p_point, neg, p_big)
        how = str(rng.choice(["inner", "left", "right", "outer"]))
        closed = bool(rng.random() < 0.3)
        ev1, ev2, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, nb, how=how, closed=closed)
        same((ev1, ev2), oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how=how, closed=closed), f"overlap {how} closed={closed}", cfg)
        if how == "left" and not closed:  # the same left join born in (row1, row2) order
            a, b, _ = engine.overlap_ordered(g1, s1, e1, g2, s2, e2, nb)
            w1, w2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how="left")
            o = np.lexsort([np.where(w2 < 0, n2 + 1, w2), w1])
            same((a, b), (w1[o], w2[o]), "overlap_ordered", cfg)
        cnt = engine.count_overlaps(g1, s1, e1, g2, s2, e2, nb, closed=closed)
        i, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how="inner", closed=closed)
        same((cnt,), (np.bincount(i, minlength=n1),), "count_overlaps", cfg)
        k = int(rng.choice([1, 2, 3, 7, 40]))
        kw = dict(k=k, ignore_overlaps=bool(rng.random() < 0.3), ignore_upstream=bool(rng.random() < 0.2),
                  ignore_downstream=bool(rng.random() < 0.2))
        along = (rng.random(n1) < 0.5) if rng.random() < 0.4 else None
        got = engine.closest_intidxs(g1, s1, e1, g2, s2, e2, nb, along=along, **kw)
        same(got[:2], oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, along=along, **kw), f"closest {kw}", cfg)
        if n1 > 1 and not kw["ignore_overlaps"]:
            got = engine.closest_intidxs(g1, s1, e1, None, None, None, nb, self_closest=True, **kw)
            same(got[:2], oo.closest_intidxs_coded(g1, s1, e1, None, None, None, nb, self_closest=True, **kw), f"self closest {kw}", cfg)
        if not p_big:  # cluster / coverage need the real ends (INT64_MAX ends only for overlap-like ops)
            md = rng.choice([0, None, 1, 50]) if True else 0
            md = None if md is None else int(md)
            res = engine.cluster(g1, s1, e1, nb, min_dist=md, want_row_spans=True)
            ids, cs, ce, cn = oo.cluster_coded(g1, s1, e1, nb, md)
            same((res["ids"], res["cstart"], res["cend"], res["ccount"]), (ids, cs, ce, cn), f"cluster {md}", cfg)
            cov = engine.coverage(g1, s1, e1, g2, s2, e2, nb)
            same((cov,), (oo.coverage_coded(g1, s1, e1, g2, s2, e2, nb),), "coverage", cfg)
            got = engine.subtract(g1, s1, e1, g2, s2, e2, nb)
            same(got, oo.subtract_coded(g1, s1, e1, g2, s2, e2, nb), "subtract", cfg)
        if n1 and n2 and len(ev1):
            lo, hi, dist = engine.pair_coords(ev1, ev2, s1, e1, s2, e2, want_distance=True)
            a, b = ev1.cpu().numpy(), ev2.cpu().numpy()
            ok = (a >= 0) & (b >= 0)
            with np.errstate(over="ignore"):
                want = (np.where(ok, np.maximum(s1[a], s2[b]), 0), np.where(ok, np.minimum(e1[a], e2[b]), 0),
                        np.where(ok, np.maximum(0, np.maximum(s1[a] - e2[b], s2[b] - e1[a])), 0))
            same((lo, hi, dist), want, "pair_coords", cfg)
    return it


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    it = run(budget, seed)
    print(f"fuzz ok: {it} rounds in {budget:.0f} s (seed {seed})")


if __name__ == "__main__":
    main()
