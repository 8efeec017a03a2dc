"""Column-name defaults and column checks: the host-side mirror of
`bioframe.core.specs` (reference src/bioframe/core/specs.py:11-152), kept so the
glue resolves column names and raises exactly as the reference does."""
from __future__ import annotations

from collections.abc import Iterable, Mapping

import numpy as np
import pandas as pd

__all__ = ["is_chrom_dtype", "update_default_colnames"]

_rc = {"colnames": {"chrom": "chrom", "start": "start", "end": "end"}}


def _get_default_colnames():
    """(chrom, start, end) names currently in force (specs.py:14-26)."""
    names = _rc["colnames"]
    return names["chrom"], names["start"], names["end"]


class update_default_colnames:
    """Set the default names; usable as a context manager (specs.py:29-58)."""

    def __init__(self, new_colnames):
        self._saved = dict(_rc["colnames"])
        bad = ValueError("Please, specify new columns using a list of 3 strings or a dict!")
        # NOTE: like the reference, a Mapping is also Iterable and takes the first branch.
        if isinstance(new_colnames, Iterable):
            if len(new_colnames) != 3:
                raise bad
            for slot, name in zip(("chrom", "start", "end"), new_colnames):
                _rc["colnames"][slot] = name
        elif isinstance(new_colnames, Mapping):
            for slot in ("chrom", "start", "end"):
                if slot in new_colnames:
                    _rc["colnames"][slot] = new_colnames[slot]
        else:
            raise bad

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        _rc["colnames"] = self._saved


def _verify_columns(df, colnames, unique_cols=False, return_as_bool=False):
    """ValueError (or False) unless `df` is a DataFrame holding `colnames` (specs.py:61-92)."""
    if not isinstance(df, pd.DataFrame):
        if return_as_bool:
            return False
        raise ValueError("df is not a dataframe")
    if unique_cols and len(set(colnames)) < len(colnames):
        raise ValueError("column names must be unique")
    missing = set(colnames).difference(df.columns)
    if missing:
        if return_as_bool:
            return False
        raise ValueError(", ".join(missing) + " not in keys of df.columns")
    return True if return_as_bool else None


def is_chrom_dtype(chrom_dtype):
    """object, string or categorical (specs.py:142-152)."""
    t = pd.api.types
    return bool(
        t.is_string_dtype(chrom_dtype)
        or t.is_object_dtype(chrom_dtype)
        or isinstance(chrom_dtype, pd.CategoricalDtype)
    )


def _verify_column_dtypes(df, cols=None, return_as_bool=False):
    """TypeError (or False) unless chrom is string-like and start/end are integer (specs.py:95-139)."""
    ck, sk, ek = _get_default_colnames() if cols is None else cols
    if not _verify_columns(df, [ck, sk, ek], return_as_bool=True):
        if return_as_bool:
            return False
        raise ValueError("could not verify columns")
    problems = []
    if not is_chrom_dtype(df.dtypes[ck]):
        problems.append("invalid df['chrom'] dtype, must be object, string, or categorical")
    if not pd.api.types.is_integer_dtype(df.dtypes[sk]):
        problems.append("invalid df['start'] dtype, must be integer")
    if not pd.api.types.is_integer_dtype(df.dtypes[ek]):
        problems.append("invalid df['end'] dtype, must be integer")
    if problems:
        if return_as_bool:
            return False
        raise TypeError(problems[0])
    return True if return_as_bool else None
