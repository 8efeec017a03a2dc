"""View construction for `complement`: the host-side mirror of
`bioframe.core.construction.make_viewframe` (reference
src/bioframe/core/construction.py:189-262) for the input kinds the six ops
use -- {chrom: length} dicts, Series of lengths, lists of (chrom, start, end
[, name]) tuples and ready DataFrames -- with the validation of
`checks.is_viewframe` (checks.py:180-250).  Views are tiny (one row per
region), so this is plain pandas on the host."""
from __future__ import annotations

import numpy as np
import pandas as pd

from .core import checks
from .core.specs import _get_default_colnames, _verify_columns


def _from_any(regions, name_col, cols):
    ck, sk, ek = cols
    if isinstance(regions, pd.DataFrame):
        if not {ck, sk, ek}.issubset(regions.columns):
            raise ValueError("Unknown dataFrame format: check column names")
        return regions.copy()
    if isinstance(regions, dict):
        rows = []
        for chrom, length in dict(regions).items():
            if not np.isscalar(length):
                raise ValueError("Unsupported dict format: {type(v)}")
            rows.append([chrom, 0, length])
        return pd.DataFrame(rows, columns=[ck, sk, ek])
    if isinstance(regions, pd.Series):
        return pd.DataFrame({ck: regions.index.values, sk: 0, ek: regions.values})
    if isinstance(regions, (list, tuple)):
        items = [regions] if np.shape(regions) == (3,) else list(regions)
        if len(np.shape(regions)) == 1 and len(regions) and isinstance(regions[0], str) and np.shape(regions) != (3,):
            raise ValueError("UCSC region strings are not supported here: pass (chrom, start, end) tuples")
        frame = pd.DataFrame(items)
        if frame.shape[1] == 3:
            frame.columns = [ck, sk, ek]
        elif frame.shape[1] == 4:
            frame.columns = [ck, sk, ek, name_col]
        else:
            raise ValueError("wrong number of columns for list input format")
        return frame
    raise ValueError(f"Unknown input format: {type(regions)}")


def _regions_overlap(view_df, cols):
    """True if two regions on one chromosome share a base (merge would join them)."""
    ck, sk, ek = cols
    ordered = view_df.sort_values([ck, sk, ek], kind="stable")
    same = ordered[ck].values[1:] == ordered[ck].values[:-1]
    reach = ordered.groupby(ck, observed=True, sort=False)[ek].cummax().values[:-1]
    return bool(np.any(same & (ordered[sk].values[1:] < reach)))


def make_viewframe(regions, view_name_col="name", cols=None):
    cols = list(_get_default_colnames() if cols is None else cols)
    ck = cols[0]
    view_df = _from_any(regions, view_name_col, cols)
    if view_name_col not in view_df.columns:
        view_df[view_name_col] = view_df[ck].values
    if not _verify_columns(view_df, [*cols, view_name_col], return_as_bool=True):
        raise TypeError("Invalid view: invalid column names")
    if not checks.is_bedframe(view_df, cols=cols):
        raise ValueError("Invalid view: not a bedframe")
    if pd.isna(view_df).values.any():
        raise ValueError("Invalid view: cannot contain NAs")
    if len(set(view_df[view_name_col])) < len(view_df):
        raise ValueError("Invalid view: entries in region_df[view_name_col] must be unique")
    if _regions_overlap(view_df, cols):
        raise ValueError("Invalid view: entries must be non-overlapping")
    return view_df
