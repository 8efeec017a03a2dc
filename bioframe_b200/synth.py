"""Deterministic synthetic hg38-shaped interval sets.

Used by bench.py, the oracle timing leg and the parity tests alike, so the
CUDA path and the CPU oracle always see the same rows.  Chromosome lengths are
the 24 primary hg38 chromosomes (the rows of the reference's
``src/bioframe/io/data/hg38.seqinfo.tsv:2-25``).  Rows are left unsorted.
"""
from __future__ import annotations

import numpy as np

HG38_CHROMS = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
HG38_SIZES = np.array(
    [
        248956422, 242193529, 198295559, 190214555, 181538259, 170805979,
        159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
        114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895, 57227415,
    ],
    dtype=np.int64,
)


def subset_fraction(chrom_ids):
    """Share of the genome the given chromosome indices cover (for bounded samples)."""
    return float(HG38_SIZES[list(chrom_ids)].sum() / HG38_SIZES.sum())


def _chrom_codes(rng, n, chrom_ids=None):
    if chrom_ids is not None:
        ids = np.asarray(list(chrom_ids), dtype=np.int32)
        w = HG38_SIZES[ids] / HG38_SIZES[ids].sum()
        cdf = np.cumsum(w)
        cdf[-1] = 1.0
        return ids[np.searchsorted(cdf, rng.random(n), side="right")]
    p = HG38_SIZES / HG38_SIZES.sum()
    # inverse-CDF draw: much cheaper than rng.choice(p=...) at 1e8 rows
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(n), side="right").astype(np.int32)


def _place(rng, codes, lens):
    span = HG38_SIZES[codes] - lens
    np.maximum(span, 1, out=span)
    starts = np.floor(rng.random(codes.shape[0]) * span).astype(np.int64)
    return starts, starts + lens


def lengths(rng, n, kind):
    """Interval lengths for one of the BASELINE.json config families."""
    if kind == "geom1k":  # cfg0 / cfg1: Geometric(1/1000), mean 1 kb
        return rng.geometric(1.0 / 1000.0, n).astype(np.int64)
    if kind == "geom2k":  # cfg3 queries
        return rng.geometric(1.0 / 2000.0, n).astype(np.int64)
    if kind == "pareto":  # cfg2 heavy tail, x_m=200, alpha=1.2, capped at 5 Mb
        x = np.ceil(200.0 * (1.0 + rng.pareto(1.2, n)))
        return np.minimum(x, 5_000_000).astype(np.int64)
    if kind == "peaks":  # cfg3 targets: LogNormal(ln 500, 0.8)
        x = np.ceil(rng.lognormal(np.log(500.0), 0.8, n))
        return np.maximum(x, 1).astype(np.int64)
    raise ValueError(kind)


def make_intervals(n, seed, kind="geom1k", chrom_ids=None):
    """-> (chrom_code int32[n], start int64[n], end int64[n]); unsorted rows.
    `chrom_ids` restricts the draw to a subset of chromosomes (bounded samples
    of a workload keep its density by scaling n with `subset_fraction`)."""
    rng = np.random.default_rng(seed)
    codes = _chrom_codes(rng, n, chrom_ids)
    lens = lengths(rng, n, kind)
    starts, ends = _place(rng, codes, lens)
    return codes, starts, ends


def make_peak_targets(n, seed, hotspot_frac=0.7, n_hotspots=200_000, window=20_000, chrom_ids=None):
    """cfg3 targets: peak-like lengths, `hotspot_frac` of them inside
    `n_hotspots` windows of `window` bp (fan-out skew)."""
    rng = np.random.default_rng(seed)
    lens = lengths(rng, n, "peaks")
    n_hot = int(n * hotspot_frac)
    hs_codes = _chrom_codes(rng, n_hotspots, chrom_ids)
    hs_starts, _ = _place(rng, hs_codes, np.full(n_hotspots, window, dtype=np.int64))
    pick = rng.integers(0, n_hotspots, n_hot)
    codes_hot = hs_codes[pick]
    starts_hot = hs_starts[pick] + np.floor(rng.random(n_hot) * window).astype(np.int64)
    ends_hot = np.minimum(starts_hot + lens[:n_hot], HG38_SIZES[codes_hot])
    starts_hot = np.minimum(starts_hot, ends_hot - 1)
    codes_bg = _chrom_codes(rng, n - n_hot, chrom_ids)
    starts_bg, ends_bg = _place(rng, codes_bg, lens[n_hot:])
    codes = np.concatenate([codes_hot, codes_bg])
    starts = np.concatenate([starts_hot, starts_bg])
    ends = np.concatenate([ends_hot, ends_bg])
    perm = rng.permutation(n)
    return codes[perm], starts[perm], ends[perm]


def to_frame(codes, starts, ends, categorical=False):
    import pandas as pd

    names = np.array(HG38_CHROMS, dtype=object)
    if categorical:
        chrom = pd.Categorical.from_codes(codes, categories=HG38_CHROMS)
    else:
        chrom = names[codes]
    return pd.DataFrame({"chrom": chrom, "start": starts, "end": ends})


def make_tie_free(n, seed, kind="geom1k", chrom_ids=None):
    """`make_intervals` with (chrom, start) and (chrom, end) both unique: rows that collide are drawn again.
    For `closest`, whose choice among targets with equal ends / equal starts is undefined in the reference
    (unstable argsort, arrops.py:562, 580; SURVEY.md §8c ii)."""
    codes, starts, ends = make_intervals(n, seed, kind, chrom_ids)
    rng = np.random.default_rng(seed + 7919)
    for _ in range(64):
        ks = codes.astype(np.int64) << 40 | starts
        ke = codes.astype(np.int64) << 40 | ends
        bad = np.zeros(n, dtype=bool)
        for key in (ks, ke):
            order = np.argsort(key, kind="stable")
            dup = np.zeros(n, dtype=bool)
            dup[order[1:]] = key[order[1:]] == key[order[:-1]]
            bad |= dup
        m = int(bad.sum())
        if m == 0:
            return codes, starts, ends
        lens = ends[bad] - starts[bad]
        s, e = _place(rng, codes[bad], lens)
        starts[bad], ends[bad] = s, e
    raise RuntimeError("could not draw a tie-free set")


HG38_OFFSETS = np.concatenate([[0], np.cumsum(HG38_SIZES)]).astype(np.int64)  # chromosomes laid end to end
HG38_TOTAL = int(HG38_OFFSETS[-1])


def linear_to_chrom(pos):
    """position on the end-to-end axis -> (chromosome code, coordinate on it); HG38_TOTAL -> (24, 0)."""
    pos = np.asarray(pos, dtype=np.int64)
    code = np.searchsorted(HG38_OFFSETS, pos, side="right") - 1
    code = np.minimum(code, len(HG38_SIZES))
    return code.astype(np.int32), pos - HG38_OFFSETS[np.minimum(code, len(HG38_SIZES) - 1)] * (code < len(HG38_SIZES))


def range_bounds(rank, world):
    """(g_lo, start_lo, g_hi, start_hi) of shard `rank` when the end-to-end genome axis is cut into `world`
    equal (chrom, start) ranges (the query sharding of SURVEY.md 8e)."""
    lo, hi = HG38_TOTAL * rank // world, HG38_TOTAL * (rank + 1) // world
    (g_lo,), (s_lo,) = (x.reshape(1) for x in linear_to_chrom(lo))
    if rank + 1 >= world:
        return int(g_lo), int(s_lo), len(HG38_SIZES), 0
    (g_hi,), (s_hi,) = (x.reshape(1) for x in linear_to_chrom(hi))
    return int(g_lo), int(s_lo), int(g_hi), int(s_hi)


def make_intervals_in_range(n, seed, kind, rank, world, sort=False):
    """n intervals whose (chrom, start) lies in shard `rank`'s range of `range_bounds`: starts uniform on that
    stretch of the end-to-end axis, lengths of `kind`, ends clipped to the chromosome.  sort=True orders the rows
    by (chrom, start) (cfg4: "queries pre-sorted by (chrom, start) on the host")."""
    rng = np.random.default_rng(seed)
    lo, hi = HG38_TOTAL * rank // world, HG38_TOTAL * (rank + 1) // world
    pos = lo + np.floor(rng.random(n) * (hi - lo)).astype(np.int64)
    if sort:
        pos.sort()
    lens = lengths(rng, n, kind)
    codes, starts = linear_to_chrom(pos)
    ends = np.minimum(starts + lens, HG38_SIZES[codes])
    return codes, starts, ends
