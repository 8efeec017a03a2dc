"""`bioframe_b200.core.arrops` -- the array-level drop-in the north star names
(reference src/bioframe/core/arrops.py:290-503, 601-754; docs/api-lowlevel.md:76-97) -- on every case of
tests/golden/arrops_cases.npz (outputs of the reference itself), twice: with the oracle-backed engine double
(CPU, checks the mirror's own logic: the [:n_matched] slice, the lone-row split, sort=True) and, marked gpu,
through the C ABI on the CUDA engine."""
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import has_cuda, reference_paths
from test_oracle_golden import CLOSEST_ALONG, CLOSEST_VARIANTS, MERGE_TAGS, eq

BIG = np.iinfo(np.int64).max


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def arrops(monkeypatch, request):
    import bioframe_b200.core.arrops as mod

    if request.param == "oracle":
        import oracle_engine

        monkeypatch.setattr(mod, "_engine", oracle_engine)
        monkeypatch.setattr(mod, "_host", lambda t: np.asarray(t))
    else:
        from bioframe_b200 import engine

        assert has_cuda() and mod._engine is engine
    return mod


def test_overlap_merge_complement_cases(arrops, arrops_golden, request):
    g = arrops_golden
    on_cuda = "cuda" in request.node.name
    for name in g["names"]:
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        if name == "wide" and on_cuda:
            # one group whose STARTS span -100 .. 2^63-6: no 64-bit (position, length) key holds it; the CUDA
            # engine refuses loudly instead of answering wrongly
            with pytest.raises(OverflowError):
                arrops.overlap_intervals(s1, e1, s2, e2)
            continue
        for closed in (0, 1):
            i, j = arrops.overlap_intervals(s1, e1, s2, e2, closed=bool(closed))
            assert i.dtype == np.int64 and j.dtype == np.int64
            eq(i, g[f"{name}/ov{closed}_1"]), eq(j, g[f"{name}/ov{closed}_2"])
            # sort=True: the same pairs ordered by (id1, id2)  (arrops.py:370-373)
            si, sj = arrops.overlap_intervals(s1, e1, s2, e2, closed=bool(closed), sort=True)
            order = np.lexsort([g[f"{name}/ov{closed}_2"], g[f"{name}/ov{closed}_1"]])
            eq(si, g[f"{name}/ov{closed}_1"][order]), eq(sj, g[f"{name}/ov{closed}_2"][order])
        i, j, u1, u2 = arrops.overlap_intervals_outer(s1, e1, s2, e2)
        eq(i, g[f"{name}/ov0_1"]), eq(j, g[f"{name}/ov0_2"])
        eq(u1, g[f"{name}/outer_u1"]), eq(u2, g[f"{name}/outer_u2"])
        if len(s1) == 0:
            continue
        for tag, md in MERGE_TAGS:
            big_end = bool(len(e1)) and int(np.max(e1)) == BIG
            if big_end and md:  # numpy wrap-around of running_max + min_dist: not reproduced
                continue
            ids, cs, ce = arrops.merge_intervals(s1, e1, md)
            eq(ids, g[f"{name}/merge_{tag}_ids"])
            eq(cs, g[f"{name}/merge_{tag}_s"]), eq(ce, g[f"{name}/merge_{tag}_e"])
        for tag in ("def", "b", "out"):
            b = tuple(int(x) for x in g[f"{name}/compl_{tag}_bounds"])
            cs, ce = arrops.complement_intervals(s1, e1, bounds=b)
            eq(cs, g[f"{name}/compl_{tag}_s"]), eq(ce, g[f"{name}/compl_{tag}_e"])


def test_closest_cases(arrops, arrops_golden):
    g = arrops_golden
    for name in g["closest_names"]:
        s1, e1, s2, e2 = (g[f"{name}/{k}"] for k in ("s1", "e1", "s2", "e2"))
        along = g[f"{name}/along"]
        for tag, kw in CLOSEST_VARIANTS.items():
            i, j = arrops.closest_intervals(s1, e1, s2, e2, **kw)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])
        for tag, kw in CLOSEST_ALONG.items():
            i, j = arrops.closest_intervals(s1, e1, s2, e2, along=along, **kw)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])
        for tag, k in (("self_k1", 1), ("self_k3", 3)):
            i, j = arrops.closest_intervals(s2, e2, None, None, k=k)
            eq(i, g[f"{name}/{tag}_1"]), eq(j, g[f"{name}/{tag}_2"])


def test_api_lowlevel_doc_example(arrops):
    """The arrays of docs/api-lowlevel.md:46-56 and the answers SURVEY.md §8c recorded from the reference."""
    s1, e1 = np.array([1, 3, 8, 12]), np.array([5, 8, 10, 14])
    s2, e2 = np.array([4, 10]), np.array([8, 11])
    i, j = arrops.overlap_intervals(s1, e1, s2, e2)
    eq(i, [0, 1]), eq(j, [0, 0])
    i, j, u1, u2 = arrops.overlap_intervals_outer(s1, e1, s2, e2)
    eq(i, [0, 1]), eq(j, [0, 0]), eq(u1, [2, 3]), eq(u2, [1])
    ids, cs, ce = arrops.merge_intervals(s1, e1, min_dist=None)
    eq(ids, [0, 0, 1, 2]), eq(cs, [1, 8, 12]), eq(ce, [8, 10, 14])
    cs, ce = arrops.complement_intervals(s1, e1)
    eq(cs, [0, 10, 14]), eq(ce, [1, 12, BIG])
    i, j = arrops.closest_intervals(s1, e1, s2, e2, k=2, ignore_downstream=True)
    eq(i, [0, 1, 2, 3, 3]), eq(j, [0, 0, 0, 1, 0])
    i, j = arrops.closest_intervals(s1, e1, k=2)
    eq(j, [1, 2, 0, 2, 1, 3, 2, 1])


def test_empty_inputs_and_errors(arrops):
    z = np.array([], dtype=np.int64)
    s, e = np.array([3, 7]), np.array([5, 9])
    for a, b, c, d in ((z, z, z, z), (s, e, z, z), (z, z, s, e)):
        i, j = arrops.overlap_intervals(a, b, c, d)
        assert i.dtype == np.int64 and len(i) == 0 and len(j) == 0
    i, j, u1, u2 = arrops.overlap_intervals_outer(s, e, z, z)
    assert len(i) == 0 and len(u2) == 0
    eq(u1, [0, 1])
    i, j, u1, u2 = arrops.overlap_intervals_outer(z, z, s, e)
    eq(u2, [0, 1])
    ids, cs, ce = arrops.merge_intervals(z, z)
    assert len(ids) == 0 and len(cs) == 0 and len(ce) == 0
    cs, ce = arrops.complement_intervals(z, z)  # numpy >= 2: the reference returns the bounds (probed live below)
    eq(cs, [0]), eq(ce, [BIG])
    i, j = arrops.closest_intervals(s, e, z, z)
    assert len(i) == 0 and len(j) == 0
    with pytest.raises(ValueError):
        arrops.merge_intervals(s, e, min_dist=-1)
    with pytest.raises(ValueError):  # the reference's lexsort error (arrops.py:738-740)
        arrops.closest_intervals(s, e, s, e, tie_arr=np.array([0, 1]))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        arrops.overlap_intervals(pd.Series(s), e, s, e)
    assert any(issubclass(x.category, SyntaxWarning) for x in w)


@pytest.mark.skipif(reference_paths()[0] is None, reason="reference not available")
def test_live_against_reference_arrops(arrops):
    """Fresh random arrays through the reference's arrops and the mirror (closest on tie-free targets)."""
    from conftest import import_reference

    import_reference()
    from bioframe.core import arrops as ref

    rng = np.random.default_rng(77)
    z = np.array([], dtype=np.int64)
    eq(arrops.complement_intervals(z, z)[1], ref.complement_intervals(z, z)[1])
    for _ in range(40):
        n1, n2 = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        s1 = rng.integers(-50, 400, n1)
        e1 = s1 + rng.integers(0, 40, n1)
        s2 = rng.integers(-50, 400, n2)
        e2 = s2 + rng.integers(0, 40, n2)
        for closed in (False, True):
            for got, want in zip(arrops.overlap_intervals(s1, e1, s2, e2, closed=closed),
                                 ref.overlap_intervals(s1, e1, s2, e2, closed=closed)):
                eq(got, want)
        for got, want in zip(arrops.overlap_intervals_outer(s1, e1, s2, e2), ref.overlap_intervals_outer(s1, e1, s2, e2)):
            eq(got, want)
        for md in (0, None, 5):
            for got, want in zip(arrops.merge_intervals(s1, e1, md), ref.merge_intervals(s1, e1, md)):
                eq(got, want)
        for b in ((0, BIG), (-100, 500), (20, 300)):
            for got, want in zip(arrops.complement_intervals(s1, e1, bounds=b), ref.complement_intervals(s1, e1, bounds=b)):
                eq(got, want)
        t2 = rng.choice(np.arange(-50, 400, 3), n2, replace=False)  # tie-free starts and ends
        te = t2 + 1
        for kw in (dict(k=1), dict(k=3), dict(k=2, ignore_overlaps=True), dict(k=2, ignore_upstream=True)):
            for got, want in zip(arrops.closest_intervals(s1, e1, t2, te, **kw), ref.closest_intervals(s1, e1, t2, te, **kw)):
                eq(got, want)
