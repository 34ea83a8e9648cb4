"""ctypes binding of ``libgiqlb200.so`` (the C ABI declared in ``include/giqlb.h``).

There is no fallback: if the shared library is missing or fails to load, importing an
operator raises ``EngineUnavailable`` telling the user how to build it.  Nothing in this
package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("GIQLB_LIB", _PKG / "_lib" / "libgiqlb200.so"))

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_RANGE, ERR_INTERNAL = -1, -2, -3, -4, -5, -6
_STATUS_NAMES = {
    ERR_INVALID: "GIQLB_ERR_INVALID", ERR_CUDA: "GIQLB_ERR_CUDA", ERR_NOMEM: "GIQLB_ERR_NOMEM",
    ERR_CAPACITY: "GIQLB_ERR_CAPACITY", ERR_RANGE: "GIQLB_ERR_RANGE", ERR_INTERNAL: "GIQLB_ERR_INTERNAL",
}

STRAND_PLUS, STRAND_MINUS, STRAND_DOT, STRAND_QMARK, STRAND_OTHER0, STRAND_NULL = 0, 1, 2, 3, 4, 255
INDEX_TIE_BY_END = 1
INDEX_NO_ROW_IDS = 2
OP_CONTAINS, OP_WITHIN = 1, 2
KEY_NULL = -2147483648   # GIQLB_KEY_NULL


class EngineUnavailable(RuntimeError):
    """The CUDA extension is not built / not loadable.  There is no CPU fallback."""


class EngineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Stats(ctypes.Structure):
    _fields_ = [
        ("rows_in", c_int64 * 2),
        ("rows_out", c_int64),
        ("algorithmic_bytes", c_int64),
        ("kernels_launched", c_int32),
        ("reserved", c_int32),
        ("kernels_total", c_int64),
    ]


class TableDesc(ctypes.Structure):
    _fields_ = [
        ("chrom_offsets", c_void_p), ("chrom_bytes", c_void_p), ("codes", c_void_p), ("codes_valid", c_int32),
        ("start", c_void_p), ("end", c_void_p), ("n", c_int64), ("start_adj", c_int32), ("end_adj", c_int32),
    ]


# name -> (restype, argtypes).  Every symbol declared in include/giqlb.h appears here; the CPU test-suite
# checks the two lists against each other and against the built library.
_P = c_void_p
PROTOTYPES = {
    "giqlb_abi_version": (c_int, []),
    "giqlb_create": (c_int, [c_int, _P, POINTER(_P)]),
    "giqlb_destroy": (c_int, [_P]),
    "giqlb_last_error": (c_char_p, [_P]),
    "giqlb_sync": (c_int, [_P]),
    "giqlb_trim": (c_int, [_P]),
    "giqlb_stats": (c_int, [_P, POINTER(Stats)]),
    "giqlb_profile_enable": (c_int, [_P, c_int]),
    "giqlb_profile_read": (c_int, [_P, POINTER(ctypes.c_double), POINTER(c_int64)]),
    "giqlb_dict_encode": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int64, _P, _P, POINTER(_P)]),
    "giqlb_dict_size": (c_int, [_P]),
    "giqlb_dict_get": (c_int, [_P, c_int32, POINTER(_P), POINTER(c_int32)]),
    "giqlb_dict_free": (c_int, [_P]),
    "giqlb_layout_create": (c_int, [_P, c_int32, c_int64, POINTER(_P)]),
    "giqlb_layout_observe": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32]),
    "giqlb_layout_finalize": (c_int, [_P, _P]),
    "giqlb_layout_free": (c_int, [_P, _P]),
    "giqlb_layout_info": (c_int, [_P, POINTER(c_int32), POINTER(c_uint64), POINTER(c_int32)]),
    "giqlb_layout_tables": (c_int, [_P, POINTER(_P), POINTER(_P), POINTER(_P)]),
    "giqlb_index_build": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int, POINTER(_P)]),
    "giqlb_index_build_rows": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int, POINTER(_P)]),
    "giqlb_index_free": (c_int, [_P, _P]),
    "giqlb_index_rows": (c_int64, [_P]),
    "giqlb_index_columns": (c_int, [_P, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P)]),
    "giqlb_index_info": (c_int, [_P, _P, POINTER(c_int64), POINTER(c_int64), POINTER(c_int32), POINTER(c_int64)]),
    "giqlb_index_group_by_strand": (c_int, [_P, _P]),
    "giqlb_intersects_plan": (c_int, [_P, _P, _P, POINTER(_P), POINTER(c_int64)]),
    "giqlb_intersects_emit": (c_int, [_P, _P, _P, _P]),
    "giqlb_joinplan_free": (c_int, [_P, _P]),
    "giqlb_intersects_join": (c_int, [_P, _P, _P, _P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_intersects_tables": (c_int, [_P, POINTER(TableDesc), POINTER(TableDesc), c_int32, _P, _P, c_int64, POINTER(c_int64),
                                        POINTER(_P)]),
    "giqlb_intersects_tables_sharded": (c_int, [_P, POINTER(TableDesc), POINTER(TableDesc), c_int32, c_int, c_int, _P, _P, c_int64,
                                                POINTER(c_int64), POINTER(_P)]),
    "giqlb_spatial_join": (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_bed_count": (c_int, [_P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_bed_parse": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P, c_int64, _P, _P, _P, _P, c_int64, _P, _P, POINTER(c_int64), POINTER(c_int64)]),
    "giqlb_pairs_filter": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int, _P, _P, c_int, _P, _P, POINTER(c_int64)]),
    "giqlb_intersects_exists": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "giqlb_select_rows": (c_int, [_P, _P, c_int64, c_int, _P, POINTER(c_int64)]),
    "giqlb_intersects_count_rows": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "giqlb_nearest": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int, c_int64, c_int, c_int,
                              _P, _P, _P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_cluster": (c_int, [_P, _P, c_int64, _P]),
    "giqlb_cluster_table": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int32, _P]),
    "giqlb_cluster_predicate": (c_int, [_P, _P, c_int64, _P, _P, c_int, _P]),
    "giqlb_merge": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_disjoin": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int, _P, _P, _P, c_int64, POINTER(c_int64)]),
    "giqlb_shard_max_end": (c_int, [_P, _P, POINTER(ctypes.c_uint32)]),
    "giqlb_cluster_shard": (c_int, [_P, _P, c_int64, ctypes.c_uint32, POINTER(ctypes.c_uint32), _P, POINTER(ctypes.c_uint32)]),
    "giqlb_merge_shard": (c_int, [_P, _P, c_int64, ctypes.c_uint32, _P, _P, _P, _P, c_int64, POINTER(c_int64), POINTER(c_int64)]),
    "giqlb_shard_cuts": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int, POINTER(ctypes.c_uint32),
                                 POINTER(ctypes.c_uint32)]),
    "giqlb_shard_select": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, ctypes.c_uint32, ctypes.c_uint32,
                                   ctypes.c_uint32, _P, POINTER(c_int64)]),
    "giqlb_index_build_range": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, ctypes.c_uint32, ctypes.c_uint32, c_int,
                                        POINTER(_P)]),
    "giqlb_distance": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int,
                               c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "giqlb_distance_rows": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int,
                                    c_int32, c_int32, c_int32, c_int32, _P, _P]),
}

PHASES = ["dict", "layout", "keygen", "sort_hist", "sort_pass", "gather", "scan_max", "join_count", "join_scan",
          "join_emit", "probe", "select", "distance", "cluster", "merge", "nearest", "other", "shard", "disjoin"]

_lib = None


def load():
    """Load the shared library once; raise EngineUnavailable (never fall back) when it cannot be."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EngineUnavailable(
            f"{LIB_PATH} not found. Build it with `make -C {_PKG / 'csrc'}` (or `python -c "
            "'import __graft_entry__ as g; g.build()'`). giql_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as exc:  # missing libcudart / wrong arch
        raise EngineUnavailable(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise EngineUnavailable(f"{LIB_PATH} does not export {name}; rebuild the extension") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib, ctx, status: int):
    if status != OK:
        msg = lib.giqlb_last_error(ctx)
        raise EngineError(status, msg.decode("utf-8", "replace") if msg else "")
