"""The CPU oracle against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py) and the hand vectors of SURVEY.md §8c.
This is what pins the oracle; the GPU tests then compare CUDA with the oracle."""
import numpy as np
import pandas as pd
import pytest

from conftest import golden_frames
from oracle import arrops_oracle as ao
from oracle import ops_oracle as oo

BIG = np.iinfo(np.int64).max

CLOSEST_VARIANTS = {
    "k1": dict(k=1),
    "k2": dict(k=2),
    "k5": dict(k=5),
    "k1_noov": dict(k=1, ignore_overlaps=True),
    "k3_noov": dict(k=3, ignore_overlaps=True),
    "k2_noup": dict(k=2, ignore_upstream=True),
    "k2_nodn": dict(k=2, ignore_downstream=True),
    "k2_noov_noup": dict(k=2, ignore_overlaps=True, ignore_upstream=True),
}
CLOSEST_ALONG = {
    "k2_along": dict(k=2),
    "k2_along_noup": dict(k=2, ignore_upstream=True),
    "k3_along_nodn_noov": dict(k=3, ignore_downstream=True, ignore_overlaps=True),
}
MERGE_TAGS = (("0", 0), ("none", None), ("2", 2), ("75", 75))


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a.astype(np.int64), b.astype(np.int64))


def test_survey_hand_vectors():
    s1, e1 = np.array([1, 3, 8, 12]), np.array([5, 8, 10, 14])
    s2, e2 = np.array([4, 10]), np.array([8, 11])
    i, j = ao.overlap_intervals(s1, e1, s2, e2)
    eq(i, [0, 1]), eq(j, [0, 0])
    i, j = ao.overlap_intervals(s1, e1, s2, e2, closed=True)
    eq(i, [0, 1, 2, 2]), eq(j, [0, 0, 1, 0])
    _, _, u1, u2 = ao.overlap_intervals_outer(s1, e1, s2, e2)
    eq(u1, [2, 3]), eq(u2, [1])
    ids, cs, ce = ao.merge_intervals(s1, e1, 0)
    eq(ids, [0, 0, 0, 1]), eq(cs, [1, 12]), eq(ce, [10, 14])
    ids, cs, ce = ao.merge_intervals(s1, e1, None)
    eq(ids, [0, 0, 1, 2]), eq(cs, [1, 8, 12]), eq(ce, [8, 10, 14])
    ids, cs, ce = ao.merge_intervals(s1, e1, 2)
    eq(ids, [0, 0, 0, 0]), eq(cs, [1]), eq(ce, [14])
    cs, ce = ao.complement_intervals(s1, e1)
    eq(cs, [0, 10, 14]), eq(ce, [1, 12, BIG])
    cs, ce = ao.complement_intervals(s1, e1, bounds=(0, 15))
    eq(ce, [1, 12, 15])
    i, j = ao.closest_intervals(s1, e1, s2, e2, k=1)
    eq(i, [0, 1, 2, 3]), eq(j, [0, 0, 0, 1])
    i, j = ao.closest_intervals(s1, e1, s2, e2, k=2)
    eq(i, [0, 0, 1, 1, 2, 2, 3, 3]), eq(j, [0, 1, 0, 1, 0, 1, 1, 0])
    i, j = ao.closest_intervals(s1, e1, s2, e2, k=1, ignore_overlaps=True)
    eq(j, [1, 1, 0, 1])
    i, j = ao.closest_intervals(s1, e1, s2, e2, k=2, ignore_overlaps=True, ignore_upstream=True)
    eq(i, [0, 1, 2]), eq(j, [1, 1, 1])
    i, j = ao.closest_intervals(s1, e1, s2, e2, k=2, ignore_downstream=True)
    eq(i, [0, 1, 2, 3, 3]), eq(j, [0, 0, 0, 1, 0])
    i, j = ao.closest_intervals(s1, e1, k=1)
    eq(j, [1, 0, 1, 2])
    i, j = ao.closest_intervals(s1, e1, k=2)
    eq(j, [1, 2, 0, 2, 1, 3, 2, 1])


def test_arrops_overlap_merge_complement(arrops_golden):
    g = arrops_golden
    for name in g["names"]:
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        for closed in (0, 1):
            i, j = ao.overlap_intervals(s1, e1, s2, e2, closed=bool(closed))
            eq(i, g[f"{name}/ov{closed}_1"]), eq(j, g[f"{name}/ov{closed}_2"])
        _, _, u1, u2 = ao.overlap_intervals_outer(s1, e1, s2, e2)
        eq(u1, g[f"{name}/outer_u1"]), eq(u2, g[f"{name}/outer_u2"])
        if len(s1) == 0:
            continue
        for tag, md in MERGE_TAGS:
            with np.errstate(over="ignore"):
                ids, cs, ce = ao.merge_intervals(s1, e1, md)
            eq(ids, g[f"{name}/merge_{tag}_ids"])
            eq(cs, g[f"{name}/merge_{tag}_s"]), eq(ce, g[f"{name}/merge_{tag}_e"])
        for tag in ("def", "b", "out"):
            b = tuple(int(x) for x in g[f"{name}/compl_{tag}_bounds"])
            cs, ce = ao.complement_intervals(s1, e1, bounds=b)
            eq(cs, g[f"{name}/compl_{tag}_s"]), eq(ce, g[f"{name}/compl_{tag}_e"])


def test_arrops_closest(arrops_golden):
    g = arrops_golden
    for name in g["closest_names"]:
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        along = g[f"{name}/along"]
        for tag, kw in CLOSEST_VARIANTS.items():
            i, j = ao.closest_intervals(s1, e1, s2, e2, **kw)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])
        for tag, kw in CLOSEST_ALONG.items():
            i, j = ao.closest_intervals(s1, e1, s2, e2, along=along, **kw)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])
        for tag, k in (("self_k1", 1), ("self_k3", 3)):
            i, j = ao.closest_intervals(s2, e2, None, None, k=k)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_frame_overlap(frame_golden, how):
    g = frame_golden
    f = golden_frames(g)
    order = list(g["overlap_group_order"])
    i, j = oo.overlap_intidxs(f["df1"], f["df2"], how=how, group_order=order)
    eq(i, g[f"overlap_{how}_1"]), eq(j, g[f"overlap_{how}_2"])


def test_frame_closest(frame_golden):
    g = frame_golden
    f = golden_frames(g)
    order = list(g["closest_group_order"])
    variants = {
        "k1": dict(k=1),
        "k3": dict(k=3),
        "k2_noov": dict(k=2, ignore_overlaps=True),
        "k2_strand_noup": dict(k=2, direction_col="strand", ignore_upstream=True),
        "k2_strand_nodn": dict(k=2, direction_col="strand", ignore_downstream=True),
    }
    for tag, kw in variants.items():
        i, j = oo.closest_intidxs(f["df1"], f["df2u"], group_order=order, **kw)
        eq(i, g[f"closest_{tag}_1"]), eq(j, g[f"closest_{tag}_2"])


def test_frame_cluster_merge(frame_golden):
    g = frame_golden
    df1 = golden_frames(g)["df1"]
    for tag, md in (("0", 0), ("none", None), ("25", 25)):
        ids, keys, cs, ce, cn = oo.cluster_core(df1, md)
        eq(ids, g[f"cluster_{tag}_ids"])
        eq(cs[ids], g[f"cluster_{tag}_cs"]), eq(ce[ids], g[f"cluster_{tag}_ce"])
        assert [k if isinstance(k, str) else k[0] for k in keys] == list(g[f"merge_{tag}_chrom"])
        eq(cs, g[f"merge_{tag}_s"]), eq(ce, g[f"merge_{tag}_e"]), eq(cn, g[f"merge_{tag}_n"])


def test_frame_coverage(frame_golden):
    g = frame_golden
    f = golden_frames(g)
    eq(oo.coverage_core(f["df1"], f["df2"]), g["coverage"])
