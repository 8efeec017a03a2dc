"""bioframe_b200: B200-native engine behind bioframe's interval-operation API.

    import bioframe_b200 as bf
    bf.overlap(df1, df2, how="inner")      # same call as bioframe.overlap
    bf.core.arrops.overlap_intervals(...)  # same call as bioframe.core.arrops.overlap_intervals

The six frame-level operations (`ops.py`) and the array-level functions
(`core/arrops.py`) keep the reference's names, arguments, errors and results;
the arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI
(`include/bioframe_b200.h`).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import core, extras  # noqa: F401
from .core.specs import update_default_colnames  # noqa: F401
from .extras import mark_runs, merge_runs, pair_by_distance  # noqa: F401
from .ops import (  # noqa: F401
    assign_view, closest, cluster, complement, count_overlaps, coverage, merge, overlap, setdiff, subtract, trim,
)

__all__ = ["overlap", "closest", "cluster", "merge", "complement", "coverage", "count_overlaps", "setdiff", "subtract",
           "assign_view", "trim", "pair_by_distance", "mark_runs", "merge_runs", "core", "extras", "update_default_colnames"]
