"""CUDA overlap path against the CPU oracle, through the C ABI (engine.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ops_oracle as oo  # noqa: E402


def _oracle_frame(g1, s1, e1, g2, s2, e2, nb, how):
    return oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how)


def _rand(rng, n, nb, span, maxlen, p_point=0.1, neg=False):
    g = rng.integers(0, nb, n).astype(np.int32)
    s = rng.integers(-span if neg else 0, span, n).astype(np.int64)
    ln = rng.integers(1, maxlen, n).astype(np.int64)
    ln[rng.random(n) < p_point] = 0
    return g, s, s + ln


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
@pytest.mark.parametrize("n1,n2,nb,span", [(0, 0, 1, 10), (5, 0, 3, 50), (0, 7, 3, 50), (1, 1, 1, 5),
                                            (300, 200, 1, 500), (5000, 3000, 7, 4000),
                                            (20000, 9000, 25, 100000), (9000, 30000, 300, 3000)])
def test_overlap_matches_oracle(how, n1, n2, nb, span):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 * 7 + n2 * 3 + nb)
    g1, s1, e1 = _rand(rng, n1, nb, span, 60, neg=(nb == 7))
    g2, s2, e2 = _rand(rng, n2, nb, span, 60, neg=(nb == 7))
    ev1, ev2, n_pairs = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, nb, how=how)
    o1, o2 = _oracle_frame(g1, s1, e1, g2, s2, e2, nb, how)
    assert ev1.shape[0] == o1.shape[0]
    assert np.array_equal(ev1.cpu().numpy(), o1)
    assert np.array_equal(ev2.cpu().numpy(), o2)
    assert n_pairs == int(((o1 >= 0) & (o2 >= 0)).sum())


def test_overlap_closed_and_golden(arrops_golden):
    from bioframe_b200 import engine

    g = arrops_golden
    for name in g["names"]:
        if name == "wide":
            continue  # needs the wide-coordinate path
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        for closed in (0, 1):
            ev1, ev2, _ = engine.overlap_intidxs(None, s1, e1, None, s2, e2, 1, how="inner", closed=bool(closed))
            assert np.array_equal(ev1.cpu().numpy(), g[f"{name}/ov{closed}_1"])
            assert np.array_equal(ev2.cpu().numpy(), g[f"{name}/ov{closed}_2"])


def test_overlap_medium_synthetic():
    """cfg0-shaped (scaled down) hg38 synthetic, all four joins."""
    from bioframe_b200 import engine, synth

    g1, s1, e1 = synth.make_intervals(200_000, 1)
    g2, s2, e2 = synth.make_intervals(150_000, 2)
    s1, e1, s2, e2 = s1 // 50, e1 // 50 + 1, s2 // 50, e2 // 50 + 1
    for how in ("inner", "outer"):
        ev1, ev2, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 24, how=how)
        o1, o2 = _oracle_frame(g1, s1, e1, g2, s2, e2, 24, how)
        assert np.array_equal(ev1.cpu().numpy(), o1)
        assert np.array_equal(ev2.cpu().numpy(), o2)


@pytest.mark.parametrize("how", ["inner", "outer"])
def test_overlap_int64_max_ends(how):
    """complement / subtract / trim feed INT64_MAX ends into overlap (ops.py:1309,1499,1604)"""
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    rng = np.random.default_rng(9)
    g1 = np.arange(5, dtype=np.int32)
    s1 = np.zeros(5, dtype=np.int64)
    e1 = np.full(5, big, dtype=np.int64)  # a view: one whole-chromosome region per group
    g2, s2, e2 = _rand(rng, 4000, 5, 10**9, 5000)
    e2[::7] = big
    s2[::11] -= 10**6  # negative starts are legal
    ev1, ev2, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 5, how=how)
    o1, o2 = _oracle_frame(g1, s1, e1, g2, s2, e2, 5, how)
    assert np.array_equal(ev1.cpu().numpy(), o1)
    assert np.array_equal(ev2.cpu().numpy(), o2)


def test_overlap_clamped_tie_order():
    """rows sharing a start whose ends all lie beyond the largest start: the clamp (INT64_MAX present)
    must not change their (end, row) order"""
    from bioframe_b200 import engine

    big = np.iinfo(np.int64).max
    s1 = np.array([10, 10, 10, 10, 3, 3, 10], dtype=np.int64)
    e1 = np.array([big, 500, 400, big - 5, 900, 800, 450], dtype=np.int64)
    s2 = np.array([0, 5, 12, 12], dtype=np.int64)
    e2 = np.array([big, 20, 13, big - 1], dtype=np.int64)
    for how in ("inner", "outer"):
        ev1, ev2, _ = engine.overlap_intidxs(None, s1, e1, None, s2, e2, 1, how=how)
        o1, o2 = _oracle_frame(None, s1, e1, None, s2, e2, 1, how)
        assert np.array_equal(ev1.cpu().numpy(), o1) and np.array_equal(ev2.cpu().numpy(), o2)


def test_overlap_row_offsets():
    from bioframe_b200 import engine

    rng = np.random.default_rng(21)
    g1, s1, e1 = _rand(rng, 3000, 4, 5000, 50)
    g2, s2, e2 = _rand(rng, 2000, 4, 5000, 50)
    a0, b0, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 4, how="outer")
    a1, b1, _ = engine.overlap_intidxs(g1, s1, e1, g2, s2, e2, 4, how="outer", row_offset1=10**10, row_offset2=7)
    a0, b0, a1, b1 = (t.cpu().numpy() for t in (a0, b0, a1, b1))
    assert np.array_equal(a1, np.where(a0 >= 0, a0 + 10**10, -1))
    assert np.array_equal(b1, np.where(b0 >= 0, b0 + 7, -1))


@pytest.mark.parametrize("n", [0, 1, 5000, 300000])
def test_sort_pairs(n):
    from bioframe_b200 import engine

    rng = np.random.default_rng(n)
    a = rng.integers(-1, 50000, n).astype(np.int64)
    b = rng.integers(-1, 7000, n).astype(np.int64)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    engine.sort_pairs(ta, tb, 49999, 6999)
    ka = np.where(a < 0, 50000, a)
    kb = np.where(b < 0, 7000, b)
    order = np.lexsort([kb, ka])
    assert np.array_equal(ta.cpu().numpy(), a[order]) and np.array_equal(tb.cpu().numpy(), b[order])


def test_pair_coords_match_numpy():
    """bf_pair_coords = the gathers of return_overlap / return_distance (ops.py:484-492, 1207-1221)."""
    from bioframe_b200 import engine

    rng = np.random.default_rng(8)
    n1, n2, n = 5000, 3000, 200_000
    s1 = rng.integers(-50, 10**6, n1)
    e1 = s1 + rng.integers(0, 500, n1)
    s2 = rng.integers(-50, 10**6, n2)
    e2 = s2 + rng.integers(0, 500, n2)
    a = rng.integers(-1, n1, n)
    b = rng.integers(-1, n2, n)
    lo, hi, dist = engine.pair_coords(a, b, s1, e1, s2, e2, want_overlap=True, want_distance=True)
    ok = (a >= 0) & (b >= 0)
    assert np.array_equal(lo.cpu().numpy(), np.where(ok, np.maximum(s1[a], s2[b]), 0))
    assert np.array_equal(hi.cpu().numpy(), np.where(ok, np.minimum(e1[a], e2[b]), 0))
    want = np.maximum(np.maximum(0, s1[a] - e2[b]), np.maximum(0, s2[b] - e1[a]))
    assert np.array_equal(dist.cpu().numpy(), np.where(ok, want, 0))
    lo, hi, dist = engine.pair_coords(a[:0], b[:0], s1, e1, s2, e2, want_distance=True)
    assert lo.numel() == 0 and dist.numel() == 0
    only = engine.pair_coords(a, b, s1, e1, s2, e2, want_overlap=False, want_distance=True)
    assert only[0] is None and only[1] is None and only[2].numel() == n


@pytest.mark.parametrize("n", [0, 1, 63, 64, 5000, (1 << 24) + 777, 40_000_001])
def test_download_rows_matches_plain_copy(n):
    """bf_download_rows: int32 over PCIe + host widening gives the same int64 host array as a plain copy."""
    from bioframe_b200 import engine

    g = torch.Generator(device="cuda").manual_seed(n)
    t = torch.randint(-1, 2**31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    if n > 2:
        t[0], t[1], t[-1] = -1, 2**31 - 2, -1
    got = engine.download_rows(t)
    assert got.dtype == np.int64 and np.array_equal(got, t.cpu().numpy())
    out = np.full(n + 5, 7, dtype=np.int64)
    view = engine.download_rows(t, out=out, n_threads=3)
    assert np.array_equal(view, t.cpu().numpy()) and (out[n:] == 7).all()


@pytest.mark.parametrize("mode", ["1", "0", None])
@pytest.mark.parametrize("n", [8 * 4096 + 1, 100_003, 1_000_000])
def test_download_rows_direct_lane_into_pinned_destination(n, mode, monkeypatch):
    """A page-locked destination lets bf_download_rows send the tail of the array as plain int64 copies on a second
    stream while the host threads widen the front (include/bioframe_b200.h).  Forced here (BF_DOWNLOAD_DIRECT=1)
    with small chunks so that the lanes meet many times in different places; =0 / unset must give the same array."""
    from bioframe_b200 import engine

    monkeypatch.setattr(engine._DownloadStage, "CHUNK_ROWS", 4096)
    monkeypatch.setattr(engine._DownloadStage, "_cache", {})
    if mode is None:
        monkeypatch.delenv("BF_DOWNLOAD_DIRECT", raising=False)
    else:
        monkeypatch.setenv("BF_DOWNLOAD_DIRECT", mode)
    g = torch.Generator(device="cuda").manual_seed(n)
    t = torch.randint(-1, 2**31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    want = t.cpu().numpy()
    out = engine.pinned_rows(n + 3)
    for threads in (1, 3):
        out[:] = 7
        got = engine.download_rows(t, out=out, n_threads=threads)
        assert np.array_equal(got, want) and (out[n:] == 7).all()
        direct = engine.download_direct_rows()
        if mode == "1":
            assert 0 < direct < n  # both lanes carried rows
        elif mode == "0":
            assert direct == 0
        else:
            assert 0 <= direct < n
    # a pageable destination never takes the second lane
    plain = np.empty(n, dtype=np.int64)
    assert np.array_equal(engine.download_rows(t, out=plain), want) and engine.download_direct_rows() == 0


def test_download_rows_rejects_values_outside_int32():
    from bioframe_b200 import _lib, engine

    t = torch.arange(100_000, dtype=torch.int64, device="cuda")
    t[777] = 2**31 - 1
    with pytest.raises(OverflowError):
        engine.download_rows(t)
    t[777] = -2
    with pytest.raises(OverflowError):
        engine.download_rows(t)
    assert _lib is not None


@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.float64, np.float32])
def test_gather_rows_matches_numpy_take(dtype):
    """bf_gather_rows = df.iloc[events] for one fixed-width column (ops.py:499-508); -1 gives 0 / invalid."""
    from bioframe_b200 import engine

    rng = np.random.default_rng(3)
    col = (rng.normal(0, 1e6, 70_001)).astype(dtype)
    idx = rng.integers(-1, len(col), 300_000)
    got, valid = engine.gather_rows(col, idx, want_valid=True)
    want = np.where(idx >= 0, col[np.maximum(idx, 0)], dtype(0)).astype(dtype)
    assert got.cpu().numpy().dtype == np.dtype(dtype)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(valid.cpu().numpy(), (idx >= 0).astype(np.uint8))
    assert engine.gather_rows(col, idx[:0]).numel() == 0


def _ordered_reference(g1, s1, e1, g2, s2, e2, nb):
    ev1, ev2 = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, nb, how="left")
    key2 = np.where(ev2 < 0, len(s2) + 1, ev2)
    order = np.lexsort([key2, ev1])
    return ev1[order], ev2[order]


@pytest.mark.parametrize("n1,n2,nb,span,maxlen,big", [
    (0, 10, 2, 100, 10, 0.0), (10, 0, 2, 100, 10, 0.0), (1, 1, 1, 10, 5, 0.0), (300, 200, 1, 500, 30, 0.0),
    (5000, 3000, 7, 4000, 60, 0.0), (20000, 9000, 25, 100000, 400, 0.0), (9000, 30000, 300, 3000, 20, 0.0),
    (4000, 2500, 3, 800, 400, 0.0), (3000, 2000, 4, 5000, 50, 0.02), (2000, 60000, 2, 300, 40, 0.0),
    # axes beyond 2^32: the 64-bit target arrays and the block-skipping walk instead of the packed records
    (6000, 4000, 2, 10**12, 5000, 0.0), (3000, 5000, 5, 10**10, 10**7, 0.0),
])
def test_overlap_ordered_matches_sorted_left_join(n1, n2, nb, span, maxlen, big):
    """bf_overlap_ordered_*: how='left' rows in (row1, row2) order = the reference's keep_order=True result
    (ops.py:549-550) at the index level; includes lists above 32 partners (heap-sorted) and INT64_MAX ends."""
    from bioframe_b200 import engine

    rng = np.random.default_rng(n1 + 7 * n2 + nb)

    def make(n):
        g = rng.integers(0, nb, n).astype(np.int32)
        s = rng.integers(0, span, n).astype(np.int64)
        ln = rng.integers(1, maxlen, n).astype(np.int64)
        ln[rng.random(n) < 0.1] = 0
        e = s + ln
        if big:
            e[rng.random(n) < big] = np.iinfo(np.int64).max
        return g, s, e

    g1, s1, e1 = make(n1)
    g2, s2, e2 = make(n2)
    ev1, ev2, biggest = engine.overlap_ordered(g1, s1, e1, g2, s2, e2, nb)
    want1, want2 = _ordered_reference(g1, s1, e1, g2, s2, e2, nb)
    assert np.array_equal(ev1.cpu().numpy(), want1)
    assert np.array_equal(ev2.cpu().numpy(), want2)
    paired = want1[want2 >= 0]
    assert biggest == (int(np.bincount(paired).max()) if len(paired) else 0)
