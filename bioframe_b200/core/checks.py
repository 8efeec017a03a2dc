"""Entry validation of the six ops: the host-side mirror of
`bioframe.core.checks.is_bedframe` (reference src/bioframe/core/checks.py:20-87)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from .specs import _get_default_colnames, _verify_column_dtypes, _verify_columns

__all__ = ["is_bedframe", "is_cataloged"]


def is_bedframe(df, raise_errors=False, cols=None):
    """chrom/start/end present with valid dtypes, nulls only as whole-null rows,
    and start <= end everywhere.  Same messages and exception classes as the
    reference (TypeError for names/dtypes, ValueError for nulls and order)."""
    ck, sk, ek = _get_default_colnames() if cols is None else cols

    def fail(exc):
        if raise_errors:
            raise exc
        return False

    if not _verify_columns(df, [ck, sk, ek], return_as_bool=True):
        return fail(TypeError("Invalid bedFrame: Invalid column names"))
    if not _verify_column_dtypes(df, cols=[ck, sk, ek], return_as_bool=True):
        return fail(TypeError("Invalid bedFrame: Invalid column dtypes"))
    plain_coords = all(isinstance(df[c].dtype, np.dtype) and df[c].dtype.kind in "iu" for c in (sk, ek))
    if plain_coords:
        # plain integer coordinates cannot be null: a null chromosome then IS a partially null row (same verdict as
        # the row-wise test below, without building a 3-column boolean frame: 50 ms per 1e7 rows)
        partial = df[ck].isna()
    else:
        nulls = pd.isnull(df[[ck, sk, ek]])
        partial = nulls.any(axis=1) & ~nulls.all(axis=1)
    if partial.any():
        return fail(
            ValueError(
                "Invalid bedFrame: Invalid null values "
                "(if any of chrom, start, end are null, then all must be null)"
            )
        )
    backwards = (df[ek].to_numpy() < df[sk].to_numpy()) if plain_coords else ((df[ek] - df[sk]) < 0)
    if backwards.any():
        return fail(
            ValueError(f"Invalid bedframe: starts exceed ends for {sum(backwards)} intervals")
        )
    return True


def is_cataloged(df, view_df, raise_errors=False, df_view_col="view_region", view_name_col="name"):
    """Every (non-null) region name in df[df_view_col] is a name of the view (checks.py:88-142)."""

    def fail(message):
        if raise_errors:
            raise ValueError(message)
        return False

    if not _verify_columns(df, [df_view_col], return_as_bool=True):
        return fail(f"Could not find `{df_view_col}` column in df")
    if not _verify_columns(view_df, [view_name_col], return_as_bool=True):
        return fail(f"Could not find `{view_name_col}` column in view_df")
    known = set(view_df[view_name_col].values)
    used = set(df[df_view_col].dropna().values)
    if not used.issubset(known):
        missing = set(df[df_view_col].values).difference(known)
        return fail(f"The following regions in df[df_view_col] not in view_df[view_name_col]: \n{missing}")
    return True
