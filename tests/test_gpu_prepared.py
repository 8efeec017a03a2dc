"""Prepared sets and (group, start)-range shards on the CUDA engine (bf_extent_measure, bf_overlap_prepare_set,
bf_overlap_adopt_set, bf_overlap_count_prepared with own_range, bf_overlap_set_halo_rows, bf_overlap_group_counts):
the shards of one join are run one after the other on one GPU -- the same calls a multi-GPU job makes, one rank
per shard -- and the assembled rows must equal the oracle's result for the whole join, row for row."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_range_sharding import NB, _sets  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def up(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(dev()) for a in arrays]


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_prepared_equals_one_shot(how):
    from bioframe_b200 import engine
    from oracle import ops_oracle as oo

    q, t = _sets(11, n1=20000, n2=6000)
    dq, dt = up(*q), up(*t)
    w1, p1 = engine.measure_extent(*dq, NB)
    w2, p2 = engine.measure_extent(*dt, NB)
    assert p1 and not p2  # _sets sorts the queries
    words = np.maximum(w1, w2)
    set1 = engine.prepare_set(words, *dq, NB, presorted=p1)
    set2 = engine.prepare_set(words, *dt, NB, presorted=p2)
    ev1, ev2, n_pairs, counts = engine.overlap_prepared(set1, set2, how=how, want_group_counts=True)
    o1, o2 = oo.overlap_intidxs_coded(*q, *t, NB, how=how)
    assert np.array_equal(ev1.cpu().numpy(), o1) and np.array_equal(ev2.cpu().numpy(), o2)
    assert n_pairs == int(((o1 >= 0) & (o2 >= 0)).sum())
    assert int(counts[:, :2].sum()) == n_pairs and int(counts.sum()) == len(o1)
    # a set sorted elsewhere and adopted (what a rank does with the broadcast keys + ids) gives the same rows,
    # and a prepared set serves any number of calls
    set2b = engine.adopt_set(words, set2.n, NB, grp=dt[0])
    set2b.keys.copy_(set2.keys)
    set2b.ids.copy_(set2.ids)
    for _ in range(2):
        a, b, _ = engine.overlap_prepared(set1, set2b, how=how)
        assert np.array_equal(a.cpu().numpy(), o1) and np.array_equal(b.cpu().numpy(), o2)


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("how", ["inner", "left"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_range_shards_assemble_to_the_reference_order(world, how, seed):
    from bioframe_b200 import engine, sharded
    from oracle import ops_oracle as oo

    q, t = _sets(seed, n1=30000, n2=5000, span=40000)
    shards = sharded.plan_range_shards(q[0], q[1], world, *t, NB)
    dt = up(*t)
    parts, counts = [], []
    for sh in shards:
        rows = np.concatenate([np.arange(sh["lo"], sh["hi"]), sh["halo"]])
        dq = up(q[0][rows], q[1][rows], q[2][rows])
        ev1, ev2, n_pairs, c = sharded.overlap_range_sharded(
            *dq, sh["hi"] - sh["lo"], sh["own_range"], *dt, NB, how=how, row_offset=sh["lo"],
            halo_ids=torch.from_numpy(sh["halo"]).to(dev()), engine=engine)
        parts.append((ev1.cpu().numpy(), ev2.cpu().numpy()))
        counts.append(c)
    a, b = sharded.assemble_host(parts, np.stack(counts))
    w1, w2 = oo.overlap_intidxs_coded(*q, *t, NB, how=how)
    assert np.array_equal(a, w1) and np.array_equal(b, w2)
    assert any(len(sh["halo"]) for sh in shards)


def test_range_shards_unsorted_inside_the_shard():
    """Rows of a shard need not be sorted (the weak-scaling bench draws them at random inside the rank's range)."""
    from bioframe_b200 import engine, sharded
    from oracle import ops_oracle as oo

    q, t = _sets(4, n1=40000, n2=8000, span=60000)
    shards = sharded.plan_range_shards(q[0], q[1], 3, *t, NB)
    rng = np.random.default_rng(0)
    perm_all = np.arange(len(q[1]))
    for sh in shards:  # shuffle the rows inside every shard: the whole set stays range-partitioned
        perm_all[sh["lo"]:sh["hi"]] = sh["lo"] + rng.permutation(sh["hi"] - sh["lo"])
    q = tuple(x[perm_all] for x in q)
    dt = up(*t)
    parts, counts = [], []
    for r, sh in enumerate(shards):
        g_hi, s_hi = sh["own_range"][2], sh["own_range"][3]
        reach = sharded.halo_reach(*t, (g_hi, s_hi), NB)
        later = np.arange(sh["hi"], len(q[1]))
        halo = later[(q[0][later] == g_hi) & (q[1][later] < reach)]
        rows = np.concatenate([np.arange(sh["lo"], sh["hi"]), halo])
        dq = up(q[0][rows], q[1][rows], q[2][rows])
        ev1, ev2, _, c = sharded.overlap_range_sharded(
            *dq, sh["hi"] - sh["lo"], sh["own_range"], *dt, NB, how="left", row_offset=sh["lo"],
            halo_ids=torch.from_numpy(halo).to(dev()), engine=engine)
        parts.append((ev1.cpu().numpy(), ev2.cpu().numpy()))
        counts.append(c)
    a, b = sharded.assemble_host(parts, np.stack(counts))
    w1, w2 = oo.overlap_intidxs_coded(*q, *t, NB, how="left")
    assert np.array_equal(a, w1) and np.array_equal(b, w2)
