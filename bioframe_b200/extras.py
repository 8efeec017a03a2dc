"""Operations the reference layers on top of `overlap` (src/bioframe/extras.py) -- SURVEY.md 8f "next" #3.
Only the layer whose cost is the pair enumeration is here: `pair_by_distance` (extras.py:389-543)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import ops
from .core.specs import _get_default_colnames

__all__ = ["pair_by_distance"]


def pair_by_distance(df, min_sep, max_sep, min_intervening=None, max_intervening=None, relative_to="midpoints",
                     cols=None, return_index=False, keep_order=False, suffixes=("_1", "_2")):
    """All pairs of intervals of `df` whose separation lies in [min_sep, max_sep] bp, optionally limited by the
    number of intervals between them (extras.py:389-543).  Every interval throws a probe to its right and one
    to its left; a pair qualifies when the right probe of the upstream interval meets the left probe of the
    downstream one -- an inner overlap of the two probe sets, enumerated on the device."""
    df = df.copy()
    stem = return_index if isinstance(return_index, str) else "index"
    icol1, icol2 = stem + suffixes[0], stem + suffixes[1]
    if return_index or keep_order:
        df.index.name = stem
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    df = df.sort_values([ck, sk, ek]).reset_index(drop=False)

    if min_sep >= max_sep:
        raise ValueError("min_sep must be less than max_sep")
    if min_sep < 0:
        raise ValueError("min_sep must be >=0")
    if min_intervening is None:
        min_intervening = 0
    if max_intervening is None:
        max_intervening = df.index.max()
    if min_intervening > max_intervening:
        raise ValueError("min_intervening must be less or equal to max_intervening")
    if min_intervening < 0:
        raise ValueError("min_intervening must be >=0")
    if relative_to not in ("midpoints", "endpoints"):
        raise ValueError("relative_to must either specify 'midpoints' or 'endpoints' ")

    mids = (df[sk] + df[ek]) // 2
    anchor_right = df[ek] if relative_to == "endpoints" else mids  # where the probe to the right starts from
    anchor_left = df[sk] if relative_to == "endpoints" else mids
    carried = [ck, stem] if (return_index or keep_order) else [ck]
    right_probe = df[carried].copy()
    right_probe[sk] = anchor_right + min_sep // 2
    right_probe[ek] = anchor_right + (max_sep + 1) // 2
    left_probe = df[carried].copy()
    left_probe[sk] = anchor_left - max_sep // 2
    left_probe[ek] = anchor_left - (min_sep + 1) // 2

    pairs = ops.overlap(right_probe, left_probe, suffixes=suffixes, how="inner", return_index=True, return_input=False)
    between = np.abs(pairs[f"index{suffixes[0]}"] - pairs[f"index{suffixes[1]}"]) - 1
    pairs = pairs[(between <= max_intervening) & (between >= min_intervening)]

    def side(suffix):
        rows = df.iloc[pairs[f"index{suffix}"].values]
        return rows.rename(columns=lambda c: c + suffix).reset_index(drop=True)

    out = pd.concat([side(suffixes[0]), side(suffixes[1])], axis=1)
    if keep_order:
        out = out.sort_values([icol1, icol2])
    if not return_index:
        out = out.drop([icol1, icol2], axis=1)
    out.reset_index(drop=True, inplace=True)
    return out
