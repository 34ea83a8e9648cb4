"""ctypes binding of ``oracle/c/giql_oracle.c`` (fast CPU restatement; TEST INFRASTRUCTURE ONLY).

Used where the numpy oracle is too slow (1e5..1e7 rows) and as the timed CPU baseline in
``bench.py`` (``cpu_baseline.kind == "port"``: it is a port of the semantics, NOT DuckDB).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libgiql_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "c" / "giql_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = ctypes.CDLL(str(_SO))
        i64, i32p, i64p, u8p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        tbl = [i64, i32p, i32p, i32p]
        _lib.giql_oracle_intersects_count.restype = i64
        _lib.giql_oracle_intersects_count.argtypes = tbl + tbl + [ctypes.c_int32, i64p, ctypes.c_int]
        _lib.giql_oracle_intersects_emit.restype = i64
        _lib.giql_oracle_intersects_emit.argtypes = tbl + tbl + [ctypes.c_int32, i64p, i32p, i32p, ctypes.c_int]
        _lib.giql_oracle_intersects_join.restype = i64
        _lib.giql_oracle_intersects_join.argtypes = tbl + tbl + [ctypes.c_int32, i64, i32p, i32p, ctypes.c_int]
        _lib.giql_oracle_intersects_exists.restype = i64
        _lib.giql_oracle_intersects_exists.argtypes = tbl + tbl + [ctypes.c_int32, u8p, ctypes.c_int]
        _lib.giql_oracle_cluster.restype = i64
        _lib.giql_oracle_cluster.argtypes = tbl + [ctypes.c_int32, i64, i32p, ctypes.c_int]
        _lib.giql_oracle_merge.restype = i64
        _lib.giql_oracle_merge.argtypes = tbl + [ctypes.c_int32, i64, i32p, i32p, i32p, i64p, ctypes.c_int]
        _lib.giql_oracle_max_threads.restype = ctypes.c_int
        _lib.giql_oracle_contains_join.restype = i64
        _lib.giql_oracle_contains_join.argtypes = tbl + tbl + [ctypes.c_int32, i64, i32p, i32p, ctypes.c_int]
        _lib.giql_oracle_nearest.restype = i64
        _lib.giql_oracle_nearest.argtypes = (tbl + [u8p]) + (tbl + [u8p]) + [ctypes.c_int32, ctypes.c_int32, i64, ctypes.c_int, ctypes.c_int,
                                                                          i32p, i32p, i64p, u8p, ctypes.c_int]
        _lib.giql_oracle_disjoin.restype = i64
        _lib.giql_oracle_disjoin.argtypes = tbl + tbl + [ctypes.c_int32, ctypes.c_int, i64p, i64p, i32p, i32p, i32p, ctypes.c_int]
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data


def max_threads() -> int:
    return int(lib().giql_oracle_max_threads())


def default_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def intersects_pairs(lc, ls, le, rc, rs, re, n_codes, threads=0, sort=True):
    """(P, 2) int64 array of (left_row, right_row), lexicographically sorted when ``sort``."""
    lc, ls, le, rc, rs, re = map(_i32, (lc, ls, le, rc, rs, re))
    counts = np.zeros(len(ls), dtype=np.int64)
    total = lib().giql_oracle_intersects_count(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re),
                                               n_codes, _p(counts), threads)
    if total < 0:
        raise MemoryError("oracle allocation failed")
    offs = np.zeros(len(ls) + 1, dtype=np.int64)
    np.cumsum(counts, out=offs[1:])
    out_l = np.empty(max(total, 1), dtype=np.int32)
    out_r = np.empty(max(total, 1), dtype=np.int32)
    lib().giql_oracle_intersects_emit(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re),
                                      n_codes, _p(offs), _p(out_l), _p(out_r), threads)
    out = np.stack([out_l[:total].astype(np.int64), out_r[:total].astype(np.int64)], axis=1)
    if sort and total:
        packed = np.sort((out[:, 0] << 32) | out[:, 1])
        out = np.stack([packed >> 32, packed & 0xFFFFFFFF], axis=1)
    return out


def intersects_join_timed(lc, ls, le, rc, rs, re, n_codes, capacity, threads=0):
    """One whole join (index build + count + scan + emit); returns (total, out_l, out_r)."""
    lc, ls, le, rc, rs, re = map(_i32, (lc, ls, le, rc, rs, re))
    out_l = np.empty(max(capacity, 1), dtype=np.int32)
    out_r = np.empty(max(capacity, 1), dtype=np.int32)
    total = lib().giql_oracle_intersects_join(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re),
                                              n_codes, capacity, _p(out_l), _p(out_r), threads)
    return int(total), out_l, out_r


def intersects_counts(lc, ls, le, rc, rs, re, n_codes, threads=0):
    lc, ls, le, rc, rs, re = map(_i32, (lc, ls, le, rc, rs, re))
    counts = np.zeros(len(ls), dtype=np.int64)
    lib().giql_oracle_intersects_count(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re),
                                       n_codes, _p(counts), threads)
    return counts


def intersects_exists(lc, ls, le, rc, rs, re, n_codes, threads=0):
    lc, ls, le, rc, rs, re = map(_i32, (lc, ls, le, rc, rs, re))
    ex = np.zeros(len(ls), dtype=np.uint8)
    lib().giql_oracle_intersects_exists(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re),
                                        n_codes, _p(ex), threads)
    return ex.astype(bool)


def cluster(part, start, end, n_parts, distance=0, threads=0):
    part, start, end = map(_i32, (part, start, end))
    ids = np.zeros(len(start), dtype=np.int32)
    lib().giql_oracle_cluster(len(start), _p(part), _p(start), _p(end), n_parts, int(distance), _p(ids), threads)
    return ids


def merge(part, start, end, n_parts, distance=0, threads=0):
    part, start, end = map(_i32, (part, start, end))
    n = len(start)
    op, os_, oe = (np.empty(max(n, 1), dtype=np.int32) for _ in range(3))
    oc = np.empty(max(n, 1), dtype=np.int64)
    m = lib().giql_oracle_merge(n, _p(part), _p(start), _p(end), n_parts, int(distance),
                                _p(op), _p(os_), _p(oe), _p(oc), threads)
    return op[:m], os_[:m], oe[:m], oc[:m]


def spatial_pairs(lc, ls, le, rc, rs, re, n_codes, op="contains", threads=0):
    """(P, 2) sorted (left_row, right_row) pairs of the CONTAINS / WITHIN column join (intersects.py:155-166)."""
    lc, ls, le, rc, rs, re = map(_i32, (lc, ls, le, rc, rs, re))
    if op == "within":                                   # l WITHIN r  <=>  r CONTAINS l
        out = spatial_pairs(rc, rs, re, lc, ls, le, n_codes, "contains", threads)
        packed = np.sort((out[:, 1] << 32) | out[:, 0])
        return np.stack([packed >> 32, packed & 0xFFFFFFFF], axis=1)
    if op != "contains":
        raise ValueError(op)
    if len(rs) and (re < rs).any():
        raise ValueError("the C oracle's CONTAINS needs end >= start on the contained side")
    total = lib().giql_oracle_contains_join(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re), n_codes, 0, None, None, threads)
    out_l, out_r = np.empty(max(total, 1), np.int32), np.empty(max(total, 1), np.int32)
    lib().giql_oracle_contains_join(len(ls), _p(lc), _p(ls), _p(le), len(rs), _p(rc), _p(rs), _p(re), n_codes, int(total), _p(out_l), _p(out_r),
                                    threads)
    packed = np.sort((out_l[:total].astype(np.int64) << 32) | out_r[:total].astype(np.int64))
    return np.stack([packed >> 32, packed & 0xFFFFFFFF], axis=1)


def nearest(qc, qs, qe, tc, ts, te, n_codes, k=1, max_distance=None, q_strand=None, t_strand=None, stranded=False, signed=False, threads=0):
    """(query_row, target_row, distance, valid) arrays of the LATERAL NEAREST form; proper intervals only.
    Strand columns are giql_b200's uint8 codes (0 '+', 1 '-', 2 '.', 3 '?', >= 4 other literals, 255 NULL)."""
    qc, qs, qe, tc, ts, te = map(_i32, (qc, qs, qe, tc, ts, te))
    if (len(qs) and (qe <= qs).any()) or (len(ts) and (te <= ts).any()):
        raise ValueError("the C oracle's NEAREST needs start < end on both sides")
    use = stranded and q_strand is not None and t_strand is not None
    qst = np.ascontiguousarray(q_strand, dtype=np.uint8) if use else None
    tst = np.ascontiguousarray(t_strand, dtype=np.uint8) if use else None
    cap = max(int(k) * len(qs), 1)
    oq, ot = np.empty(cap, np.int32), np.empty(cap, np.int32)
    od, ov = np.empty(cap, np.int64), np.empty(cap, np.uint8)
    m = lib().giql_oracle_nearest(len(qs), _p(qc), _p(qs), _p(qe), _p(qst) if use else None, len(ts), _p(tc), _p(ts), _p(te),
                                  _p(tst) if use else None, n_codes, int(k), -1 if max_distance is None else int(max_distance), int(use),
                                  int(bool(signed)), _p(oq), _p(ot), _p(od), _p(ov), threads)
    if m < 0:
        raise MemoryError("oracle allocation failed")
    return oq[:m], ot[:m], od[:m], ov[:m].astype(bool)


def disjoin(tc, ts, te, rc=None, rs=None, re=None, n_codes=0, threads=0):
    """(target_row, seg_start, seg_end) arrays sorted by (row, start); reference None = the target itself."""
    tc, ts, te = map(_i32, (tc, ts, te))
    self_ref = rc is None
    if self_ref:
        rc, rs, re = tc, ts, te
    rc, rs, re = map(_i32, (rc, rs, re))
    if len(rs) and (re < rs).any():
        raise ValueError("the C oracle's DISJOIN needs start <= end on the reference")
    counts = np.zeros(max(len(ts), 1), dtype=np.int64)
    args = (len(ts), _p(tc), _p(ts), _p(te), len(rs), _p(rc), _p(rs), _p(re), n_codes, int(self_ref))
    total = lib().giql_oracle_disjoin(*args, _p(counts), None, None, None, None, threads)
    if total < 0:
        raise MemoryError("oracle allocation failed")
    offs = np.zeros(len(ts) + 1, dtype=np.int64)
    np.cumsum(counts[: len(ts)], out=offs[1:])
    row, s, e = (np.empty(max(total, 1), np.int32) for _ in range(3))
    lib().giql_oracle_disjoin(*args, _p(counts), _p(offs), _p(row), _p(s), _p(e), threads)
    return row[:total], s[:total], e[:total]
