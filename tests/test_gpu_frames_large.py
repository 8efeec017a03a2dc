"""Frame-level calls big enough to take the large-result paths of the glue: the host download of the pair list
(bf_download_rows, results of 4 M rows and more) and the device-side clipped coordinates (bf_pair_coords)."""
import numpy as np
import pandas as pd
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _frames(n1, n2, seed):
    rng = np.random.default_rng(seed)
    chroms = np.array(["chr1", "chr2", "chr3", "chrX"], dtype=object)

    def one(n, maxlen):
        s = rng.integers(0, 3_000_000, n)
        return pd.DataFrame({"chrom": chroms[rng.integers(0, 4, n)], "start": s, "end": s + rng.integers(0, maxlen, n)})

    return one(n1, 400), one(n2, 60_000)


def test_large_overlap_frame_matches_index_level_result():
    import bioframe_b200 as bf
    from bioframe_b200 import engine, ops

    df1, df2 = _frames(300_000, 10_000, 5)
    ev1, ev2 = ops._overlap_intidxs(df1, df2, how="inner")
    assert len(ev1) >= (1 << 22), "the case must be large enough for the download path"
    # the same pairs straight from the engine, copied the plain way
    codes1, keys = pd.factorize(df1["chrom"], sort=True)
    order = list(set.union(set(keys), set(df2["chrom"].unique())))
    rank = {k: i for i, k in enumerate(order)}
    g1 = df1["chrom"].map(rank).to_numpy(np.int32)
    g2 = df2["chrom"].map(rank).to_numpy(np.int32)
    a, b, _ = engine.overlap_intidxs(g1, df1["start"].values, df1["end"].values, g2, df2["start"].values,
                                     df2["end"].values, len(order), how="inner")
    assert np.array_equal(ev1, a.cpu().numpy()) and np.array_equal(ev2, b.cpu().numpy())
    # every reported pair overlaps, and the frame carries the clipped coordinates of exactly these pairs
    s1, e1, s2, e2 = df1["start"].values, df1["end"].values, df2["start"].values, df2["end"].values
    adj = lambda s, e: e + (e == s)  # noqa: E731
    assert (s1[ev1] < adj(s2, e2)[ev2]).all() and (s2[ev2] < adj(s1, e1)[ev1]).all()
    out = bf.overlap(df1, df2, how="inner", return_overlap=True, return_index=True, return_input=False)
    assert len(out) == len(ev1)
    assert np.array_equal(out["overlap_start"].to_numpy(), np.maximum(s1[ev1], s2[ev2]))
    assert np.array_equal(out["overlap_end"].to_numpy(), np.minimum(e1[ev1], e2[ev2]))
    assert np.array_equal(out["index"].to_numpy(), ev1) and np.array_equal(out["index_"].to_numpy(), ev2)
