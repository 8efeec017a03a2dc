"""ctypes binding of libbioframe_b200.so (the C ABI in include/bioframe_b200.h).

There is no CPU fallback: if the library is missing or cannot be loaded,
`lib()` raises.  The library is built in-tree by `python -m bioframe_b200.build`
(nvcc, sm_100a) and is what `__graft_entry__.build()` produces.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# BIOFRAME_B200_LIB: another build of the same sources (the range-checked one, build.py --checked); never a fallback
LIB_PATH = os.environ.get("BIOFRAME_B200_LIB") or os.path.join(_PKG, "libbioframe_b200.so")

BF_OK = 0
ERRORS = {
    -1: "BF_ERR_ARG",
    -2: "BF_ERR_WORKSPACE",
    -3: "BF_ERR_RANGE",
    -4: "BF_ERR_CUDA",
    -5: "BF_ERR_GROUP",
    -6: "BF_ERR_STATE",
}
HOW = {"inner": 0, "left": 1, "right": 2, "outer": 3}
CLOSEST_IGNORE_OVERLAPS, CLOSEST_IGNORE_UPSTREAM, CLOSEST_IGNORE_DOWNSTREAM, CLOSEST_SELF = 1, 2, 4, 8


class Plan(C.Structure):
    _fields_ = [("opaque", C.c_int64 * 96)]


class SortedSetStruct(C.Structure):
    _fields_ = [("opaque", C.c_int64 * 32)]


class EngineError(RuntimeError):
    pass


_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/bioframe_b200.h declares
SIGNATURES = {
    "bf_last_error": (C.c_char_p, []),
    "bf_version": (C.c_int, []),
    "bf_checked_build": (C.c_int, []),
    "bf_prof_enable": (C.c_int, [C.c_int]),
    "bf_prof_reset": (C.c_int, []),
    "bf_prof_num_classes": (C.c_int, []),
    "bf_prof_class_name": (C.c_char_p, [C.c_int]),
    "bf_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "bf_overlap_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "bf_overlap_count": (
        C.c_int,
        [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _sz, C.POINTER(Plan), _p,
         C.POINTER(_i64), C.POINTER(_i64)],
    ),
    "bf_overlap_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p]),
    "bf_overlap_set_row_offsets": (C.c_int, [C.POINTER(Plan), _i64, _i64]),
    "bf_extent_words": (_sz, [_i32]),
    "bf_extent_workspace_bytes": (_sz, [_i32]),
    "bf_extent_measure": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _sz, _p, _p, C.POINTER(_i32)]),
    "bf_sorted_set_bytes": (_sz, [_i64, _i32, _i32]),
    "bf_overlap_prepare_set": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _sz, _p, C.POINTER(SortedSetStruct)]),
    "bf_overlap_adopt_set": (C.c_int, [_p, _i64, _i32, _p, _sz, _p, C.POINTER(SortedSetStruct)]),
    "bf_sorted_set_buffers": (C.c_int, [C.POINTER(SortedSetStruct), C.POINTER(_p), C.POINTER(_p), C.POINTER(_i64)]),
    "bf_sorted_set_runmax": (C.c_int, [C.POINTER(SortedSetStruct), _p]),
    "bf_overlap_prepared_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "bf_overlap_count_prepared": (
        C.c_int,
        [C.POINTER(SortedSetStruct), C.POINTER(SortedSetStruct), _p, _p, _i32, _i32, _p, _p, _sz, C.POINTER(Plan), _p,
         C.POINTER(_i64), C.POINTER(_i64)],
    ),
    "bf_overlap_set_halo_rows": (C.c_int, [C.POINTER(Plan), _i64, _p]),
    "bf_overlap_group_counts": (C.c_int, [C.POINTER(Plan), _p, _p, _p]),
    "bf_overlap_ordered_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bf_overlap_ordered_count": (
        C.c_int,
        [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p, _sz, C.POINTER(Plan), _p, C.POINTER(_i64), C.POINTER(_i64)],
    ),
    "bf_overlap_ordered_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p]),
    "bf_sort_pairs_workspace_bytes": (_sz, [_i64]),
    "bf_sort_pairs": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _sz, _p]),
    "bf_gather_rows": (C.c_int, [_p, _i32, _p, _i64, _p, _p, _p]),
    "bf_group_codes_present": (C.c_int, [_p, _i32, _i64, _i32, _p, _p]),
    "bf_group_codes_apply": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p]),
    "bf_pair_coords": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "bf_closest_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bf_closest_count": (
        C.c_int,
        [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _sz, C.POINTER(Plan), _p,
         C.POINTER(_i64), C.POINTER(_i64)],
    ),
    "bf_closest_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p]),
    "bf_cluster_workspace_bytes": (_sz, [_i64, _i32]),
    "bf_cluster_count": (
        C.c_int,
        [_p, _p, _p, _i64, _i32, _i64, _i32, _p, _sz, C.POINTER(Plan), _p, _p, C.POINTER(_i64)],
    ),
    "bf_cluster_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "bf_coverage_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bf_coverage": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p, _sz, _p, _p]),
    "bf_count_overlaps_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bf_count_overlaps": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _i32, _p, _sz, _p, _p]),
    "bf_complement_workspace_bytes": (_sz, [_i64, _i32]),
    "bf_complement_count": (
        C.c_int,
        [_p, _p, _p, _i64, _p, _p, _i32, _p, _sz, C.POINTER(Plan), _p, C.POINTER(_i64)],
    ),
    "bf_complement_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p, _p]),
    "bf_download_stage_bytes": (_sz, [_i64]),
    "bf_download_rows": (C.c_int, [_p, _i64, _p, _p, _sz, _p, _sz, _i32, _p]),
    "bf_download_direct_rows": (_i64, []),
    "bf_bed_scan_workspace_bytes": (_sz, [_i64]),
    "bf_bed_scan": (C.c_int, [_p, _i64, _p, _sz, _p, C.POINTER(_i64)]),
    "bf_bed_parse_workspace_bytes": (_sz, [_i64, _i64]),
    "bf_bed_parse": (
        C.c_int,
        [_p, _i64, _i64, _p, _sz, C.POINTER(Plan), _p, C.POINTER(_i64), C.POINTER(_i32), _p, _p, _p],
    ),
    "bf_bed_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p, _p, _p]),
    "bf_subtract_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bf_subtract_count": (
        C.c_int,
        [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p, _sz, C.POINTER(Plan), _p, C.POINTER(_i64)],
    ),
    "bf_subtract_emit": (C.c_int, [C.POINTER(Plan), _p, _p, _p, _p, _p, _p]),
}

_lib = None


def lib():
    """Load the shared library once; fail loudly if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} not found: build it with `python -m bioframe_b200.build` "
                "(there is no CPU fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != BF_OK:
        msg = lib().bf_last_error().decode("utf-8", "replace")
        kind = ERRORS.get(rc, str(rc))
        if rc in (-1, -5):
            raise ValueError(f"{what}: {kind}: {msg}")
        if rc == -3:
            raise OverflowError(f"{what}: {kind}: {msg}")
        raise EngineError(f"{what}: {kind}: {msg}")


def prof_enable(on=True):
    lib().bf_prof_enable(1 if on else 0)


def prof_reset():
    lib().bf_prof_reset()


def prof_read():
    """-> {class_name: dict(ms, timed, launches, bytes)} for classes that launched."""
    L = lib()
    out = {}
    for c in range(L.bf_prof_num_classes()):
        ms, timed, launches, nbytes = C.c_double(0), _i64(0), _i64(0), _i64(0)
        L.bf_prof_read(c, C.byref(ms), C.byref(timed), C.byref(launches), C.byref(nbytes))
        if launches.value:
            out[L.bf_prof_class_name(c).decode()] = dict(
                ms=ms.value, timed=timed.value, launches=launches.value, bytes=nbytes.value
            )
    return out
