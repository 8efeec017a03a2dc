"""Live differential test against the reference itself.  Random frames go through the reference's public
functions and through this repo's frame-level functions -- every case twice: on the oracle-backed test double
(CPU: the reference's pandas/numpy pipeline against oracle + glue) and, marked gpu, on the CUDA engine (GPU box,
reference = the vendored git-ignored copy of oracle/make_ref.py).  Complements the frozen goldens of
tests/golden/ with fresh draws."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import has_cuda, import_reference, reference_paths

pytestmark = pytest.mark.skipif(reference_paths()[0] is None, reason="reference checkout / vendored copy not present")


@pytest.fixture(scope="module")
def ref():
    return import_reference()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def ours(monkeypatch, request):
    import bioframe_b200 as bf
    import bioframe_b200.core.arrops as arrops
    import bioframe_b200.ops as ops

    if request.param == "oracle":
        import oracle_engine

        monkeypatch.setattr(ops, "_engine", oracle_engine)
        monkeypatch.setattr(arrops, "_engine", oracle_engine)
    else:
        from bioframe_b200 import engine

        assert has_cuda() and ops._engine is engine and arrops._engine is engine
    old = pd.get_option("future.infer_string")
    pd.set_option("future.infer_string", False)
    yield bf
    pd.set_option("future.infer_string", old)


def _frame(rng, n, chroms=("chr1", "chr2", "chrX"), span=600, maxlen=50, p_point=0.1):
    s = rng.integers(0, span, n)
    ln = rng.integers(1, maxlen, n)
    ln[rng.random(n) < p_point] = 0
    return pd.DataFrame({"chrom": rng.choice(np.array(chroms, dtype=object), n), "start": s, "end": s + ln,
                         "score": rng.integers(0, 5, n)})


@pytest.mark.parametrize("seed", range(6))
def test_live_overlap_family(ref, ours, seed):
    rng = np.random.default_rng(1000 + seed)
    df1, df2 = _frame(rng, int(rng.integers(1, 120))), _frame(rng, int(rng.integers(1, 90)), chroms=("chr2", "chrX", "chr9"))
    # row order pinned by the reference: left join with keep_order, and everything per-query
    for kw in (dict(how="left"), dict(how="left", return_index=True, return_overlap=True, suffixes=("_q", "_t"))):
        pd.testing.assert_frame_equal(ours.overlap(df1, df2, **kw), ref.overlap(df1, df2, **kw))
    # chromosome blocks come in set order (ops.py:289, 301): the same process builds the same sets, so even the
    # joins without keep_order must agree row for row
    for how in ("inner", "right", "outer"):
        got, want = ours.overlap(df1, df2, how=how, return_index=True), ref.overlap(df1, df2, how=how, return_index=True)
        pd.testing.assert_frame_equal(got, want)
    got, want = ours.ops._overlap_intidxs(df1, df2, how="outer"), ref.ops._overlap_intidxs(df1, df2, how="outer")
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    pd.testing.assert_frame_equal(ours.coverage(df1.copy(), df2), ref.coverage(df1.copy(), df2))
    pd.testing.assert_frame_equal(ours.count_overlaps(df1.copy(), df2), ref.count_overlaps(df1.copy(), df2))
    pd.testing.assert_frame_equal(ours.setdiff(df1, df2), ref.setdiff(df1, df2))
    pd.testing.assert_frame_equal(ours.subtract(df1, df2, return_index=True), ref.subtract(df1, df2, return_index=True))


@pytest.mark.parametrize("seed", range(6))
def test_live_cluster_family(ref, ours, seed):
    rng = np.random.default_rng(2000 + seed)
    df = _frame(rng, int(rng.integers(1, 150)))
    for md in (0, None, 3, 40):
        pd.testing.assert_frame_equal(ours.cluster(df, min_dist=md), ref.cluster(df, min_dist=md))
        pd.testing.assert_frame_equal(ours.merge(df, min_dist=md), ref.merge(df, min_dist=md))
    pd.testing.assert_frame_equal(ours.complement(df), ref.complement(df))
    view = {"chr1": 700, "chr2": 650, "chrX": 900}
    pd.testing.assert_frame_equal(ours.complement(df, view_df=view), ref.complement(df, view_df=view))
    pd.testing.assert_frame_equal(ours.trim(df, view_df={"chr1": 300, "chr2": 400, "chrX": 200}),
                                  ref.trim(df, view_df={"chr1": 300, "chr2": 400, "chrX": 200}))
    pd.testing.assert_frame_equal(ours.mark_runs(df, "score", allow_overlaps=True), ref.mark_runs(df, "score", allow_overlaps=True))
    pd.testing.assert_frame_equal(ours.merge_runs(df, "score", allow_overlaps=True), ref.merge_runs(df, "score", allow_overlaps=True))


@pytest.mark.parametrize("forced_device_assembly", [False, True])
@pytest.mark.parametrize("seed", range(4))
def test_live_categorical_family(ref, ours, seed, forced_device_assembly, monkeypatch):
    """Categorical chromosome columns: dictionary codes become group codes without a per-row remap when the
    observed categories are the first ones (ops._factorize_column), categories out of lexical order, categories
    that never occur, whole-null rows; once more with the large-result paths (device encode + column gathers)
    forced on for these small frames."""
    import bioframe_b200.ops as ops

    if forced_device_assembly:
        monkeypatch.setattr(ops, "_DEVICE_GATHER_MIN_ROWS", 0)
        monkeypatch.setattr(ops, "_DEVICE_ENCODE_MIN_ROWS", 0)
    rng = np.random.default_rng(5000 + seed)
    cats = [["chrX", "chr2", "chr10", "chr1", "chrUn"], ["chr1", "chr2", "chrX"], ["chr9", "chrX", "chr2", "chr1"],
            ["chr2", "chr1"]][seed]
    used1, used2 = cats[: max(2, len(cats) - 1)], cats[1:]

    def cat_frame(n, used, nulls):
        df = _frame(rng, n, chroms=tuple(used))
        df["chrom"] = pd.Categorical(df["chrom"], categories=cats)
        if nulls:
            df["start"] = df["start"].astype(pd.Int64Dtype())
            df["end"] = df["end"].astype(pd.Int64Dtype())
            df.loc[df.index[3::17], ["chrom", "start", "end"]] = pd.NA
        return df

    for nulls in (False, True):
        df1, df2 = cat_frame(int(rng.integers(20, 140)), used1, nulls), cat_frame(int(rng.integers(20, 90)), used2, nulls)
        for how in ("left", "inner", "outer"):
            kw = dict(how=how, return_index=True, return_overlap=True)
            pd.testing.assert_frame_equal(ours.overlap(df1, df2, **kw), ref.overlap(df1, df2, **kw))
        for md in (0, None, 5):
            pd.testing.assert_frame_equal(ours.cluster(df1, min_dist=md), ref.cluster(df1, min_dist=md))
            pd.testing.assert_frame_equal(ours.merge(df1, min_dist=md), ref.merge(df1, min_dist=md))
        pd.testing.assert_frame_equal(ours.cluster(df2, return_cluster_ids=False), ref.cluster(df2, return_cluster_ids=False))
        pd.testing.assert_frame_equal(ours.cluster(df2, return_cluster_intervals=False, return_input=False),
                                      ref.cluster(df2, return_cluster_intervals=False, return_input=False))
        pd.testing.assert_frame_equal(ours.coverage(df1.copy(), df2), ref.coverage(df1.copy(), df2))
        pd.testing.assert_frame_equal(ours.count_overlaps(df1.copy(), df2), ref.count_overlaps(df1.copy(), df2))
        pd.testing.assert_frame_equal(ours.setdiff(df1, df2), ref.setdiff(df1, df2))
        pd.testing.assert_frame_equal(ours.subtract(df1, df2, return_index=True), ref.subtract(df1, df2, return_index=True))
        pd.testing.assert_frame_equal(ours.subtract(df1, df2), ref.subtract(df1, df2))
        # closest among equal distances goes through an unstable sort in the reference: compare what is defined
        got, want = ours.closest(df1, df2, k=1), ref.closest(df1, df2, k=1)
        assert list(got.columns) == list(want.columns) and len(got) == len(want)
        pd.testing.assert_series_equal(got["distance"].sort_values(ignore_index=True),
                                       want["distance"].sort_values(ignore_index=True))


def test_live_null_rows_under_categorical_chromosome(ref, ours):
    """`groupby([one categorical column], observed=True, dropna=False).indices` has no group for the null rows
    (ops._pandas_drops_null_category), so the reference's `_overlap_intidxs` does not return them at all -- unlike
    null rows under an object column, which come back unpaired.  Same rows here, at the index level, for the
    frame-level sharded call (one rank) and in the positional attach of count_overlaps / coverage."""
    import bioframe_b200.ops as ops
    from bioframe_b200 import sharded

    cats = ["chr2", "chr1", "chrX"]

    def frame(chroms, starts, ends):
        return pd.DataFrame({"chrom": pd.Categorical(chroms, categories=cats), "start": pd.array(starts, dtype="Int64"),
                             "end": pd.array(ends, dtype="Int64")})

    df1 = frame(["chr2", "chr2", None, "chr1", "chr1", None, "chrX"], [1, 30, None, 5, 50, None, 7], [10, 40, None, 9, 60, None, 9])
    df2 = frame(["chr2", None, "chr1", "chr1", "chrX"], [5, None, 1, 55, 100], [8, None, 6, 58, 110])
    assert ops._pandas_drops_null_category() == (len(df1.groupby(["chrom"], observed=True, dropna=False).indices) == 3)
    for how in ("left", "inner", "outer", "right"):
        got, want = ours.ops._overlap_intidxs(df1, df2, how=how), ref.ops._overlap_intidxs(df1, df2, how=how)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), how
        pd.testing.assert_frame_equal(ours.overlap(df1, df2, how=how, return_index=True, return_overlap=True),
                                      ref.overlap(df1, df2, how=how, return_index=True, return_overlap=True))
    if getattr(ops._engine, "__name__", "") == "oracle_engine":  # the sharded frame call: host logic, checked on CPU
        for how in ("left", "inner"):
            got, want = sharded.overlap_intidxs_frames(df1, df2, how=how, engine=ops._engine), ref.ops._overlap_intidxs(df1, df2, how=how)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), how
    pd.testing.assert_frame_equal(ours.count_overlaps(df1.copy(), df2), ref.count_overlaps(df1.copy(), df2))
    pd.testing.assert_frame_equal(ours.coverage(df1.copy(), df2), ref.coverage(df1.copy(), df2))
    pd.testing.assert_frame_equal(ours.setdiff(df1, df2), ref.setdiff(df1, df2))
    # the same rows under an object column stay in the join (the NaN key is a group there)
    o1, o2 = df1.astype({"chrom": object}), df2.astype({"chrom": object})
    got, want = ours.ops._overlap_intidxs(o1, o2, how="outer"), ref.ops._overlap_intidxs(o1, o2, how="outer")
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and len(got[0]) > len(df1) - 2


@pytest.mark.parametrize("seed", range(4))
def test_live_closest_tie_free(ref, ours, seed):
    """closest picks among equal ends / starts through an unstable argsort in the reference: tie-free draws."""
    rng = np.random.default_rng(3000 + seed)

    def tie_free(n, chroms):
        s = rng.choice(np.arange(0, 5000, 7), n, replace=False)
        e = s + 1 + 7 * rng.integers(0, 5, n) + rng.integers(0, 6, n)
        while len(set(e)) < n:
            dup = pd.Series(e).duplicated().values
            e[dup] = s[dup] + 1 + rng.integers(0, 80, dup.sum())
        return pd.DataFrame({"chrom": rng.choice(np.array(chroms, dtype=object), n), "start": s, "end": e})

    df1, df2 = tie_free(60, ("chr1", "chr2")), tie_free(45, ("chr2", "chr1", "chr3"))
    for kw in (dict(k=1), dict(k=3, return_overlap=True, return_index=True), dict(k=2, ignore_overlaps=True),
               dict(k=2, ignore_upstream=True), dict(k=1, ignore_downstream=True)):
        pd.testing.assert_frame_equal(ours.closest(df1, df2, **kw), ref.closest(df1, df2, **kw))  # block order too


@pytest.mark.parametrize("seed", range(3))
def test_live_nulls_in_on_column(ref, ours, seed):
    """A null in an `on` column is a key value shared by both frames (groupby(dropna=False), ops.py:287-288);
    all-null rows never pair.  Same process, same hash seed: the block order must match too."""
    rng = np.random.default_rng(4000 + seed)
    df1, df2 = _frame(rng, 80), _frame(rng, 60, chroms=("chr2", "chrX", "chr1"))
    for df in (df1, df2):
        df["strand"] = rng.choice(np.array(["+", "-", None], dtype=object), len(df))
        df["start"] = df["start"].astype(pd.Int64Dtype())
        df["end"] = df["end"].astype(pd.Int64Dtype())
        df.loc[df.index[::13], ["chrom", "start", "end"]] = pd.NA
    for how in ("inner", "left", "outer", "right"):
        kw = dict(how=how, on=["strand"], return_index=True)
        pd.testing.assert_frame_equal(ours.overlap(df1, df2, **kw), ref.overlap(df1, df2, **kw))
    pd.testing.assert_frame_equal(ours.count_overlaps(df1, df2, on=["strand"]), ref.count_overlaps(df1, df2, on=["strand"]))
    pd.testing.assert_frame_equal(ours.setdiff(df1, df2, on=["strand"]), ref.setdiff(df1, df2, on=["strand"]))


def test_block_order_when_key_hashes_collide(ref, ours):
    """The reference visits its groups in the iteration order of `set.union(set(df1_groups), set(df2_groups))`
    (ops.py:289, 301) where the two operands are built from DICTS; a set built from a list of the same keys can
    iterate differently when hashes collide in the table (CPython sizes a set from a dict once, a set from a list
    grows and re-inserts).  Integer group keys whose two constructions differ -- independent of the string hash
    seed -- must still give the reference's row order."""
    keys1, keys2 = [58, 68, 114, 178, 184], [5, 6, 7, 81, 138, 166]
    assert list(set.union(set(keys1), set(keys2))) != list(set.union(set(dict.fromkeys(keys1)), set(dict.fromkeys(keys2))))
    rng = np.random.default_rng(11)

    def frame(keys, n):
        s = rng.integers(0, 300, n)
        return pd.DataFrame({"chrom": rng.choice(np.array(keys + [58, 5]), n), "start": s, "end": s + rng.integers(1, 60, n)})

    df1, df2 = frame(keys1, 300), frame(keys2, 200)
    for how in ("inner", "left", "outer", "right"):
        a, b = ours.ops._overlap_intidxs(df1, df2, how=how)
        ra, rb = ref.ops._overlap_intidxs(df1, df2, how=how)
        assert np.array_equal(np.asarray(a), np.asarray(ra, dtype=np.int64)), how
        assert np.array_equal(np.asarray(b), np.asarray(rb, dtype=np.int64)), how
    a, b = ours.ops._closest_intidxs(df1, df2, k=2)
    ra, rb = ref.ops._closest_intidxs(df1, df2, k=2)
    assert np.array_equal(np.asarray(a), np.asarray(ra, dtype=np.int64)) and np.array_equal(np.asarray(b), np.asarray(rb, dtype=np.int64))
