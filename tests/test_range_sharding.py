"""Exact reference row order from (group, start)-range shards (bioframe_b200/sharded.py: plan_range_shards,
overlap_range_sharded, assemble_offsets).  CPU: the per-shard engine is the oracle-backed double, whose
ownership rule is an independent numpy model; the assembled result must equal the single-process oracle result
row for row.  The shards are cut so that targets straddle the boundaries (long targets, dense queries)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402

NB = 5


def _sets(seed, n1=3000, n2=700, span=4000, long_every=9):
    rng = np.random.default_rng(seed)
    g1 = rng.integers(0, NB, n1)
    s1 = rng.integers(0, span, n1)
    e1 = s1 + rng.integers(0, 60, n1)  # includes points
    order = np.lexsort([s1, g1])
    g1, s1, e1 = g1[order], s1[order], e1[order]
    g2 = rng.integers(0, NB, n2)
    s2 = rng.integers(0, span, n2)
    ln = rng.integers(0, 80, n2)
    ln[::long_every] += rng.integers(500, 3000, len(ln[::long_every]))  # targets that reach over whole shards
    return (g1.astype(np.int32), s1.astype(np.int64), e1.astype(np.int64)), (g2.astype(np.int32), s2.astype(np.int64), (s2 + ln).astype(np.int64))


def _run_shards(q, t, world, how):
    import oracle_engine
    from bioframe_b200 import sharded

    shards = sharded.plan_range_shards(q[0], q[1], world, *t, NB)
    parts, counts = [], []
    for sh in shards:
        rows = np.concatenate([np.arange(sh["lo"], sh["hi"]), sh["halo"]])
        ev1, ev2, n_pairs, c = sharded.overlap_range_sharded(
            q[0][rows], q[1][rows], q[2][rows], sh["hi"] - sh["lo"], sh["own_range"], *t, NB, how=how,
            row_offset=sh["lo"], halo_ids=sh["halo"], engine=oracle_engine)
        assert n_pairs == int((np.asarray(ev2) >= 0).sum())
        parts.append((np.asarray(ev1), np.asarray(ev2)))
        counts.append(c)
    return sharded.assemble_host(parts, np.stack(counts)), shards


@pytest.mark.parametrize("world", [1, 2, 3, 7])
@pytest.mark.parametrize("how", ["inner", "left"])
@pytest.mark.parametrize("seed", [0, 1])
def test_assembled_shards_equal_the_whole_join(world, how, seed):
    from oracle import ops_oracle as oo

    q, t = _sets(seed)
    (a, b), shards = _run_shards(q, t, world, how)
    w1, w2 = oo.overlap_intidxs_coded(*q, *t, NB, how=how)
    assert np.array_equal(a, w1) and np.array_equal(b, w2)
    if world > 2:
        assert any(len(sh["halo"]) for sh in shards)  # the case under test: targets straddle shard bounds


def test_plan_rejects_unsorted_queries_and_right_joins():
    import oracle_engine
    from bioframe_b200 import sharded

    q, t = _sets(3)
    with pytest.raises(ValueError):
        sharded.plan_range_shards(q[0][::-1], q[1][::-1], 2, *t, NB)
    with pytest.raises(ValueError):
        sharded.overlap_range_sharded(*q, len(q[1]), None, *t, NB, how="outer", engine=oracle_engine)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import oracle_engine
    from bioframe_b200 import sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, t = _sets(5)
        sh = sharded.plan_range_shards(q[0], q[1], world, *t, NB)[rank]
        rows = np.concatenate([np.arange(sh["lo"], sh["hi"]), sh["halo"]])
        tt = t if rank == 0 else (None, None, None)  # the targets exist on rank 0 only: keys + ids arrive by broadcast
        ev1, ev2, _, c = sharded.overlap_range_sharded(
            q[0][rows], q[1][rows], q[2][rows], sh["hi"] - sh["lo"], sh["own_range"], *tt, NB, how="left",
            row_offset=sh["lo"], halo_ids=sh["halo"], engine=oracle_engine, n2=len(t[1]))
        allc = sharded.gather_group_counts(c)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), ev1=np.asarray(ev1), ev2=np.asarray(ev2), counts=allc)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_broadcast_of_sorted_targets(tmp_path):
    from bioframe_b200 import sharded
    from oracle import ops_oracle as oo

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    assert np.array_equal(res[0]["counts"], res[1]["counts"])
    a, b = sharded.assemble_host([(r["ev1"], r["ev2"]) for r in res], res[0]["counts"])
    q, t = _sets(5)
    w1, w2 = oo.overlap_intidxs_coded(*q, *t, NB, how="left")
    assert np.array_equal(a, w1) and np.array_equal(b, w2)


def _frame_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import pandas as pd

    import oracle_engine
    from bioframe_b200 import sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        df1, df2 = _frames()
        a, b = sharded.overlap_intidxs_frames(df1, df2, how="left", engine=oracle_engine)
        np.savez(os.path.join(out_dir, f"f{rank}.npz"), a=a, b=b)
    finally:
        dist.destroy_process_group()


def _frames():
    import pandas as pd

    q, t = _sets(8, n1=2500, n2=600)
    names = np.array(["chr1", "chr10", "chr2", "chrX", "chrY"], dtype=object)  # sorted like the chromosome column
    df1 = pd.DataFrame({"chrom": names[q[0]], "start": q[1], "end": q[2]})
    df2 = pd.DataFrame({"chrom": names[t[0]], "start": t[1], "end": t[2]})
    return df1, df2


def test_frame_level_sharded_overlap_world2_gloo(tmp_path):
    """DataFrames in, the reference's `_overlap_intidxs` rows out, from two ranks: every rank returns the whole
    result; it must be a valid reference result for SOME visiting order of the chromosome blocks (the order follows
    the hash seed of rank 0's process), so the rows are compared block by block with the single-process glue."""
    import oracle_engine
    from bioframe_b200 import ops

    world = 2
    mp.spawn(_frame_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(tmp_path, f"f{r}.npz")) for r in range(world)]
    assert np.array_equal(res[0]["a"], res[1]["a"]) and np.array_equal(res[0]["b"], res[1]["b"])
    df1, df2 = _frames()
    saved = ops._engine
    ops._engine = oracle_engine
    try:
        w1, w2 = ops._overlap_intidxs(df1, df2, how="left")
    finally:
        ops._engine = saved
    a, b = res[0]["a"], res[0]["b"]
    assert len(a) == len(w1)
    chrom = df1["chrom"].to_numpy()
    for c in np.unique(chrom):  # the block of every chromosome, rows in the same order
        m_got, m_want = chrom[a] == c, chrom[w1] == c
        assert np.array_equal(a[m_got], w1[m_want]) and np.array_equal(b[m_got], w2[m_want]), c
