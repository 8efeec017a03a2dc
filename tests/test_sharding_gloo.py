"""N>1 host logic of the query-sharded overlap (bioframe_b200/sharded.py) with world_size 2 on CPU
(gloo).  The per-rank engine is the oracle-backed test double; what is checked is the sharding, the
target broadcast, the row offsets and that the concatenated pairs equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, how, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import oracle_engine
    from bioframe_b200 import sharded, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g1, s1, e1 = synth.make_intervals(6000, 1)
        g2, s2, e2 = synth.make_intervals(2500, 2)
        s1, e1, s2, e2 = s1 // 3000, e1 // 3000 + 1, s2 // 3000, e2 // 3000 + 1
        b = sharded.shard_bounds(len(s1), world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        # targets: real data on rank 0 only, garbage elsewhere until the broadcast
        t = [torch.from_numpy(x.copy()) for x in (g2, s2, e2)]
        if rank != 0:
            for x in t:
                x.zero_()
        ev1, ev2, off, total = sharded.overlap_sharded(
            torch.from_numpy(g1[lo:hi]), torch.from_numpy(s1[lo:hi]), torch.from_numpy(e1[lo:hi]), lo,
            t[0], t[1], t[2], 24, how=how, engine=_TensorEngine(oracle_engine))
        eng = _TensorEngine(oracle_engine)
        q = (torch.from_numpy(g1[lo:hi]), torch.from_numpy(s1[lo:hi]), torch.from_numpy(e1[lo:hi]))
        cov = sharded.coverage_sharded(*q, t[0], t[1], t[2], 24, engine=eng, broadcast=False)
        cnt = sharded.count_overlaps_sharded(*q, t[0], t[1], t[2], 24, engine=eng, broadcast=False)
        c1, c2, _ = sharded.closest_sharded(q[0], q[1], q[2], lo, t[0], t[1], t[2], 24, engine=eng, broadcast=False, k=2)
        o1, o2 = sharded.overlap_ordered_sharded(q[0], q[1], q[2], lo, t[0], t[1], t[2], 24, engine=eng, broadcast=False)
        sub = sharded.subtract_sharded(q[0], q[1], q[2], lo, t[0], t[1], t[2], 24, engine=eng, broadcast=False)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), ev1=ev1.numpy(), ev2=ev2.numpy(), off=off, total=total,
                 cov=np.asarray(cov), cnt=np.asarray(cnt), c1=np.asarray(c1), c2=np.asarray(c2),
                 o1=np.asarray(o1), o2=np.asarray(o2), sub=np.stack([np.asarray(x) for x in sub]))
    finally:
        dist.destroy_process_group()


class _TensorEngine:
    """oracle double taking torch CPU tensors"""

    def __init__(self, mod):
        self.mod = mod

    def coverage(self, g1, s1, e1, g2, s2, e2, nb):
        return self.mod.coverage(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb)

    def count_overlaps(self, g1, s1, e1, g2, s2, e2, nb):
        return self.mod.count_overlaps(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb)

    def closest_intidxs(self, g1, s1, e1, g2, s2, e2, nb, **kw):
        a, b, n = self.mod.closest_intidxs(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb, **kw)
        return torch.from_numpy(a), torch.from_numpy(b), n

    def overlap_ordered(self, g1, s1, e1, g2, s2, e2, nb):
        a, b, n = self.mod.overlap_ordered(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb)
        return torch.from_numpy(a), torch.from_numpy(b), n

    def subtract(self, g1, s1, e1, g2, s2, e2, nb):
        return tuple(torch.from_numpy(np.asarray(x)) for x in
                     self.mod.subtract(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb))

    def overlap_intidxs(self, g1, s1, e1, g2, s2, e2, nb, how="inner", closed=False, row_offset1=0, row_offset2=0):
        a, b, n = self.mod.overlap_intidxs(g1.numpy(), s1.numpy(), e1.numpy(), g2.numpy(), s2.numpy(), e2.numpy(), nb,
                                           how=how, closed=closed, row_offset1=row_offset1, row_offset2=row_offset2)
        return torch.from_numpy(a), torch.from_numpy(b), n


@pytest.mark.parametrize("how", ["inner", "left"])
def test_sharded_overlap_world2(how, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), how, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bioframe_b200 import synth
    from oracle import ops_oracle as oo

    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    assert parts[0]["off"] == 0 and parts[1]["off"] == len(parts[0]["ev1"])
    assert parts[0]["total"] == parts[1]["total"] == sum(len(p["ev1"]) for p in parts)
    ev1 = np.concatenate([p["ev1"] for p in parts])
    ev2 = np.concatenate([p["ev2"] for p in parts])
    g1, s1, e1 = synth.make_intervals(6000, 1)
    g2, s2, e2 = synth.make_intervals(2500, 2)
    s1, e1, s2, e2 = s1 // 3000, e1 // 3000 + 1, s2 // 3000, e2 // 3000 + 1
    o1, o2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 24, how=how)
    # same rows; after the (query row, target row) ordering of keep_order=True they are identical arrays
    got = np.lexsort([ev2, ev1])
    want = np.lexsort([o2, o1])
    assert np.array_equal(ev1[got], o1[want]) and np.array_equal(ev2[got], o2[want])
    # per-query results: concatenation of the shards is the single-process answer
    assert np.array_equal(np.concatenate([p["cov"] for p in parts]), oo.coverage_coded(g1, s1, e1, g2, s2, e2, 24))
    i_all, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 24, how="inner")
    assert np.array_equal(np.concatenate([p["cnt"] for p in parts]), np.bincount(i_all, minlength=len(s1)))
    # closest: same (query, target) rows as the single-process call once both are ordered by query row
    c1 = np.concatenate([p["c1"] for p in parts])
    c2 = np.concatenate([p["c2"] for p in parts])
    w1, w2 = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, 24, k=2)
    a, b = np.argsort(c1, kind="stable"), np.argsort(w1, kind="stable")
    assert np.array_equal(c1[a], w1[b]) and np.array_equal(c2[a], w2[b])
    # keep_order left join and subtract: per query, so the plain concatenation is the single-process result
    l1, l2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, 24, how="left")
    order = np.lexsort([np.where(l2 < 0, len(s2) + 1, l2), l1])
    assert np.array_equal(np.concatenate([p["o1"] for p in parts]), l1[order])
    assert np.array_equal(np.concatenate([p["o2"] for p in parts]), l2[order])
    want_sub = np.stack(oo.subtract_coded(g1, s1, e1, g2, s2, e2, 24))
    assert np.array_equal(np.concatenate([p["sub"] for p in parts], axis=1), want_sub)


def test_shard_bounds():
    from bioframe_b200 import sharded

    b = sharded.shard_bounds(10, 4)
    assert list(b) == [0, 3, 6, 8, 10]
    assert list(sharded.shard_bounds(0, 2)) == [0, 0, 0]
