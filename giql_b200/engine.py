"""Host layer of the B200 engine: ingest (numpy / pandas / Arrow / torch), device-resident tables and one
method per GIQL operator, each a thin sequence of C-ABI calls into ``libgiqlb200.so``.

PyTorch is used for device memory and stream plumbing only; every data-path kernel lives in
``giql_b200/csrc``.  There is no CPU fallback: constructing an :class:`Engine` without a CUDA device or
without the built extension raises.

Operator semantics follow the reference (abdenlab/giql), cited per method.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Any, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import EngineError, EngineUnavailable
from .table import Table, canonical_offsets

_I32_MIN, _I32_MAX = -(2**31), 2**31 - 1


def _as_i32(col, what: str) -> np.ndarray:
    """Host integer column -> contiguous int32, rejecting NULLs and out-of-range values."""
    if isinstance(col, torch.Tensor):
        col = col.detach().cpu().numpy()
    try:
        import pyarrow as pa

        if isinstance(col, (pa.Array, pa.ChunkedArray)):
            if col.null_count:
                raise ValueError(f"{what}: NULL coordinates are not supported (BED never has them)")
            col = col.to_numpy(zero_copy_only=False) if isinstance(col, pa.Array) else col.combine_chunks().to_numpy(zero_copy_only=False)
    except ImportError:  # pragma: no cover
        pass
    if hasattr(col, "to_numpy") and not isinstance(col, np.ndarray):
        col = col.to_numpy()
    a = np.asarray(col)
    if a.dtype == object:
        if any(v is None for v in a):
            raise ValueError(f"{what}: NULL coordinates are not supported (BED never has them)")
        a = a.astype(np.int64)
    if a.dtype.kind == "f":
        if np.isnan(a).any():
            raise ValueError(f"{what}: NULL coordinates are not supported (BED never has them)")
        a = a.astype(np.int64)
    if a.dtype.kind not in "iu":
        raise ValueError(f"{what}: expected an integer column, got dtype {a.dtype}")
    if a.dtype != np.int32:
        if a.size and (int(a.min()) < _I32_MIN or int(a.max()) > _I32_MAX):
            raise ValueError(f"{what}: coordinate outside the int32 range")
        a = a.astype(np.int32)
    return np.ascontiguousarray(a)


def _as_utf8(col, what: str):
    """Host string column -> (offsets int32[n+1], bytes uint8[]) in Arrow utf8 layout."""
    import pyarrow as pa

    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks()
    if not isinstance(col, pa.Array):
        if hasattr(col, "to_numpy") and not isinstance(col, np.ndarray):
            col = col.to_numpy()
        if isinstance(col, np.ndarray) and col.dtype.kind in "US":
            col = col.astype(object) if col.dtype.kind == "U" else np.array([b.decode() for b in col], dtype=object)
        col = pa.array(list(col) if not isinstance(col, np.ndarray) else col, type=pa.string(), from_pandas=True)
    if pa.types.is_dictionary(col.type):
        col = col.dictionary_decode()
    if pa.types.is_large_string(col.type) or not pa.types.is_string(col.type):
        col = col.cast(pa.string())
    if col.null_count:
        raise ValueError(f"{what}: NULL chromosome names are not supported (BED never has them)")
    n = len(col)
    bufs = col.buffers()
    offsets = np.frombuffer(bufs[1], dtype=np.int32)[col.offset: col.offset + n + 1] if n or bufs[1] is not None else np.zeros(1, np.int32)
    data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
    if n == 0:
        offsets = np.zeros(1, np.int32)
    return np.ascontiguousarray(offsets), data


class Dictionary:
    """Shared chromosome dictionary: ``names[code]`` sorted bytewise (codes preserve VARCHAR order)."""

    def __init__(self, names: Sequence[str]):
        self.names = list(names)
        self._lookup = {n: i for i, n in enumerate(self.names)}

    def __len__(self) -> int:
        return len(self.names)

    def code(self, name: str) -> int:
        return self._lookup.get(name, -1)


@dataclass
class PairResult:
    """Overlapping pairs as row ids into the two input tables (device int32 tensors, order unspecified)."""

    left_rows: torch.Tensor
    right_rows: torch.Tensor

    def __len__(self) -> int:
        return int(self.left_rows.numel())

    def sorted_packed(self) -> np.ndarray:
        """``sort((left << 32) | right)`` on the host — the reference's sorted-multiset normal form
        (tests/integration/_oracle.py:77-89)."""
        l = self.left_rows.to(torch.int64)
        r = self.right_rows.to(torch.int64)
        return torch.sort((l << 32) | r).values.cpu().numpy()

    def to_host(self):
        return self.left_rows.cpu().numpy(), self.right_rows.cpu().numpy()

    def to_arrow(self, left_name: str = "l", right_name: str = "r"):
        """pyarrow Table of the row-id pairs — the relation a SQL engine joins back on ``rowid`` to finish arbitrary
        projections / GROUP BY / ORDER BY (the `__giql_p0` / `__giql_p1` columns of intersects_duckdb.py:1339-1345)."""
        import pyarrow as pa

        l, r = self.to_host()
        return pa.table({left_name: pa.array(l, type=pa.int32()), right_name: pa.array(r, type=pa.int32())})


@dataclass
class NearestResult:
    """(query_row, target_row, distance) rows; ``valid`` is 0 where the distance is SQL NULL."""

    query_rows: torch.Tensor
    target_rows: torch.Tensor
    distance: torch.Tensor
    valid: torch.Tensor

    def __len__(self) -> int:
        return int(self.query_rows.numel())

    def rows(self):
        q, t, d, v = (x.cpu().tolist() for x in (self.query_rows, self.target_rows, self.distance, self.valid))
        return [(qi, ti, di if vi else None) for qi, ti, di, vi in zip(q, t, d, v)]


@dataclass
class DisjoinResult:
    """(target_row, segment_start, segment_end), canonical 0-based half-open coordinates."""

    target_rows: torch.Tensor
    start: torch.Tensor
    end: torch.Tensor

    def __len__(self) -> int:
        return int(self.target_rows.numel())

    def rows(self):
        return sorted(zip(self.target_rows.cpu().tolist(), self.start.cpu().tolist(), self.end.cpu().tolist()))


@dataclass
class MergeResult:
    """Merged regions (device tensors): chromosome code, start, end, optional strand code, COUNT(*)."""

    engine: Any
    dictionary: "Dictionary"
    code: torch.Tensor
    start: torch.Tensor
    end: torch.Tensor
    strand: Optional[torch.Tensor]
    count: torch.Tensor

    def __len__(self) -> int:
        return int(self.code.numel())

    def rows(self):
        """Host tuples ``(chrom, [strand,] start, end, count)`` in a canonical sorted order."""
        names = self.dictionary.names
        c, s, e, k = (x.cpu().tolist() for x in (self.code, self.start, self.end, self.count))
        if self.strand is None:
            rows = [(names[ci], si, ei, ki) for ci, si, ei, ki in zip(c, s, e, k)]
        else:
            st = [self.engine.strand_literal(v) for v in self.strand.cpu().tolist()]
            rows = [(names[ci], ti, si, ei, ki) for ci, ti, si, ei, ki in zip(c, st, s, e, k)]
        return sorted(rows, key=lambda r: tuple((v is None, v) for v in r))


class DeviceTable:
    """Columns of one genomic table resident in HBM."""

    def __init__(self, engine: "Engine", n: int, config: Optional[Table], start: torch.Tensor, end: torch.Tensor,
                 strand: Optional[torch.Tensor], chrom_offsets: Optional[torch.Tensor], chrom_bytes: Optional[torch.Tensor],
                 codes: Optional[torch.Tensor] = None, dictionary: Optional[Dictionary] = None):
        self.engine = engine
        self.n = n
        self.config = config
        self.start, self.end, self.strand = start, end, strand
        self.chrom_offsets, self.chrom_bytes = chrom_offsets, chrom_bytes
        self.codes, self.dictionary = codes, dictionary

    def __len__(self) -> int:
        return self.n

    @property
    def offsets(self) -> tuple[int, int]:
        return canonical_offsets(self.config)

    def device_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.start, self.end, self.strand, self.chrom_offsets,
                                                           self.chrom_bytes, self.codes) if t is not None)


def bind_host_to_gpu(index: int = 0):
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``index`` (same NUMA node / PCIe root), so
    that the pinned host buffers it allocates afterwards are first-touched on that node and host<->device copies do not
    cross the socket interconnect.  Returns ``(previous affinity, new affinity)``; a no-op (previous, None) when NVML is
    unavailable or the local cores are outside the cores the process may use (containers)."""
    prev = os.sched_getaffinity(0)
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {i for i in range(n_cpu) if (int(words[i // 64]) >> (i % 64)) & 1} & prev
        if not local or local == prev:
            return prev, None
        os.sched_setaffinity(0, local)
        return prev, local
    except Exception:                                               # noqa: BLE001 — affinity is an optimisation only
        return prev, None


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Layout:
    def __init__(self, engine: "Engine", n_codes: int, pad: int = 0):
        self.e = engine
        self.h = ctypes.c_void_p()
        engine._check(engine.lib.giqlb_layout_create(engine.ctx, n_codes, pad, ctypes.byref(self.h)))

    def observe(self, t: DeviceTable, raw: bool = False):
        sadj, eadj = (0, 0) if raw else t.offsets
        self.e._check(self.e.lib.giqlb_layout_observe(self.e.ctx, self.h, _ptr(t.codes), _ptr(t.start), _ptr(t.end), t.n, sadj, eadj))

    def finalize(self):
        self.e._check(self.e.lib.giqlb_layout_finalize(self.e.ctx, self.h))

    def info(self):
        n, ext, bits = ctypes.c_int32(), ctypes.c_uint64(), ctypes.c_int32()
        self.e.lib.giqlb_layout_info(self.h, ctypes.byref(n), ctypes.byref(ext), ctypes.byref(bits))
        return n.value, ext.value, bits.value

    def free(self):
        if self.h:
            self.e.lib.giqlb_layout_free(self.e.ctx, self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Index:
    def __init__(self, engine: "Engine", layout: _Layout, t: DeviceTable, raw: bool = False, with_strand: bool = False,
                 tie_by_end: bool = False, rows: Optional[torch.Tensor] = None, start_range: Optional[tuple] = None,
                 row_ids: bool = True):
        self.e = engine
        self.layout = layout   # keep the layout alive as long as the index
        self.rows = rows       # ... and the row selection until the stream has consumed it
        self.h = ctypes.c_void_p()
        sadj, eadj = (0, 0) if raw else t.offsets
        if start_range is not None:   # the rows whose linear start lies in [lo, hi): a genome-range shard, selected and keyed in one pass
            if with_strand or tie_by_end or rows is not None:
                raise ValueError("a start-range index takes no strand grouping, end ties or row selection")
            engine._check(engine.lib.giqlb_index_build_range(engine.ctx, layout.h, _ptr(t.codes), _ptr(t.start), _ptr(t.end), t.n, sadj, eadj,
                                                            int(start_range[0]), int(start_range[1]), 0 if row_ids else _cabi.INDEX_NO_ROW_IDS,
                                                            ctypes.byref(self.h)))
            return
        strand = t.strand if with_strand else None
        if with_strand and strand is None:
            raise ValueError("stranded := true requires a strand column on the table")
        flags = (_cabi.INDEX_TIE_BY_END if tie_by_end else 0) | (0 if row_ids else _cabi.INDEX_NO_ROW_IDS)
        n = t.n if rows is None else int(rows.numel())
        engine._check(engine.lib.giqlb_index_build_rows(engine.ctx, layout.h, _ptr(t.codes), _ptr(t.start), _ptr(t.end),
                                                       _ptr(strand), _ptr(rows), n, sadj, eadj, flags, ctypes.byref(self.h)))

    def info(self):
        d, ng, nb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        self.e._check(self.e.lib.giqlb_index_info(self.e.ctx, self.h, ctypes.byref(d), ctypes.byref(ng), ctypes.byref(nb), None))
        return d.value, ng.value, nb.value

    def free(self):
        if self.h:
            self.e.lib.giqlb_index_free(self.e.ctx, self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One engine per (process, device).  Work is enqueued on the torch stream current at construction."""

    def __init__(self, device: int | str | torch.device | None = None):
        self.lib = _cabi.load()   # raises EngineUnavailable when the extension is not built
        if not torch.cuda.is_available():
            raise EngineUnavailable("giql_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        self.ctx = ctypes.c_void_p()
        st = self.lib.giqlb_create(self.device.index or 0, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(self.ctx))
        if st != _cabi.OK:
            raise EngineError(st, "giqlb_create failed")
        self._strand_codes: dict[Any, int] = {"+": 0, "-": 1, ".": 2, "?": 3}
        self._pair_capacity = 0   # running estimate of the pair-buffer size for the single-pass join

    # ---- plumbing ---------------------------------------------------------------------------------
    def _check(self, status: int):
        _cabi.check(self.lib, self.ctx, status)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.giqlb_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.giqlb_sync(self.ctx))

    def trim(self):
        """Return the engine's cached scratch memory to the driver (synchronises)."""
        self._check(self.lib.giqlb_trim(self.ctx))

    def stats(self) -> dict:
        s = _cabi.Stats()
        self._check(self.lib.giqlb_stats(self.ctx, ctypes.byref(s)))
        return {"rows_in": (s.rows_in[0], s.rows_in[1]), "rows_out": s.rows_out, "algorithmic_bytes": s.algorithmic_bytes,
                "kernels_launched": s.kernels_launched, "kernels_total": s.kernels_total, "big_tiles": s.reserved}

    def profile(self, on: bool = True):
        """Record CUDA events around every kernel group (see giqlb_profile_enable)."""
        self._check(self.lib.giqlb_profile_enable(self.ctx, int(on)))

    def profile_read(self) -> dict:
        """{phase: (device ms, kernel launches)} accumulated since the previous read; synchronises."""
        ms = (ctypes.c_double * len(_cabi.PHASES))()
        ln = (ctypes.c_int64 * len(_cabi.PHASES))()
        self._check(self.lib.giqlb_profile_read(self.ctx, ms, ln))
        return {name: (ms[i], ln[i]) for i, name in enumerate(_cabi.PHASES)}

    def kernels_total(self) -> int:
        return self.stats()["kernels_total"]

    def _dev(self, a: np.ndarray, pinned: bool = False) -> torch.Tensor:
        t = torch.from_numpy(a)
        if pinned:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=pinned)

    def _empty(self, n: int, dtype) -> torch.Tensor:
        try:
            return torch.empty(max(int(n), 0), dtype=dtype, device=self.device)
        except torch.cuda.OutOfMemoryError:
            # the library parks its scratch buffers per context where torch's allocator cannot see them: hand them back
            # and retry once
            self.trim()
            torch.cuda.empty_cache()
            return torch.empty(max(int(n), 0), dtype=dtype, device=self.device)

    # ---- ingest -----------------------------------------------------------------------------------
    def strand_codes(self, col) -> np.ndarray:
        """VARCHAR strand column -> uint8 codes ('+' 0, '-' 1, '.' 2, '?' 3, other literals 4.., NULL 255).

        SQL semantics the codes preserve: ``a.strand = b.strand`` (nearest.py:328-329) is plain equality on
        non-NULL values, NULL never equals anything; PARTITION BY strand (cluster.py:243-245) groups NULLs.
        """
        import pyarrow as pa

        if isinstance(col, pa.ChunkedArray):
            col = col.combine_chunks()
        if not isinstance(col, pa.Array):
            if hasattr(col, "to_numpy") and not isinstance(col, np.ndarray):
                col = col.to_numpy()
            col = pa.array(list(col) if not isinstance(col, np.ndarray) else col.astype(object), type=pa.string(), from_pandas=True)
        if not pa.types.is_dictionary(col.type):
            col = col.dictionary_encode()
        values = col.dictionary.to_pylist()
        lut = np.empty(len(values) + 1, dtype=np.uint8)
        for i, v in enumerate(values):
            if v not in self._strand_codes:
                if len(self._strand_codes) >= 254:
                    raise ValueError("more than 250 distinct strand literals")
                self._strand_codes[v] = len(self._strand_codes)
            lut[i] = self._strand_codes[v]
        lut[len(values)] = _cabi.STRAND_NULL
        idx = col.indices.fill_null(len(values)).to_numpy(zero_copy_only=False).astype(np.int64)
        return lut[idx]

    def table(self, data, config: Table | str | None = None, *, pinned: bool = False) -> DeviceTable:
        """Upload one table.  ``data``: mapping / pandas DataFrame / pyarrow Table with the columns ``config``
        names (reference column-mapping contract: src/giql/table.py:16-93; unregistered tables default to
        chrom/start/end/strand, intersects_duckdb.py:1181-1188)."""
        if isinstance(config, str):
            config = Table(config)
        cfg = config or Table("t")
        get = (lambda k: data[k]) if not hasattr(data, "column") else (lambda k: data.column(k))
        names = set(data.keys()) if hasattr(data, "keys") else set(getattr(data, "column_names", []))
        start = _as_i32(get(cfg.start_col), cfg.start_col)
        end = _as_i32(get(cfg.end_col), cfg.end_col)
        if len(start) != len(end):
            raise ValueError("start and end columns differ in length")
        offs, byts = _as_utf8(get(cfg.chrom_col), cfg.chrom_col)
        if len(offs) - 1 != len(start):
            raise ValueError("chrom column length differs from start/end")
        strand = None
        if cfg.strand_col is not None and cfg.strand_col in names:
            strand = self._dev(self.strand_codes(get(cfg.strand_col)), pinned)
        return DeviceTable(self, len(start), config, self._dev(start, pinned), self._dev(end, pinned), strand,
                           self._dev(offs, pinned), self._dev(byts, pinned))

    # ---- BED text ingest ----------------------------------------------------------------------------------
    @staticmethod
    def _bed_data_start(head: bytes) -> int:
        """byte offset of the first data line: leading `#`, `track` and `browser` lines are header (UCSC BED format)"""
        at = 0
        while at < len(head):
            nl = head.find(b"\n", at)
            line = head[at: nl if nl >= 0 else len(head)]
            if line.startswith((b"#", b"track", b"browser")) or not line.strip():
                if nl < 0:
                    return len(head)
                at = nl + 1
                continue
            break
        return at

    def read_bed(self, source, fields: Optional[int] = None, config: Table | str | None = None) -> DeviceTable:
        """BED text -> DeviceTable, parsed on the device (giqlb_bed_count / giqlb_bed_parse).

        ``source``: a path, ``bytes`` or a uint8 numpy array with the file's content (< 2 GiB; TAB-separated, lines
        ending in \\n or \\r\\n; leading ``#`` / ``track`` / ``browser`` lines are skipped).  ``fields``: 3..6 columns to
        read (default: as many of the first six as the first data line has).  The table carries chrom / start / end and,
        from BED6, strand; ``.bed_name`` = (utf8 offsets, bytes) and ``.bed_score`` (int32) hold the other columns on
        the device.  BED is 0-based half-open, which is also the default ``Table`` encoding."""
        if isinstance(source, (str, os.PathLike)):
            raw = np.fromfile(source, dtype=np.uint8)
        elif isinstance(source, (bytes, bytearray, memoryview)):
            raw = np.frombuffer(source, dtype=np.uint8)
        else:
            raw = np.ascontiguousarray(source, dtype=np.uint8)
        skip = self._bed_data_start(raw[: 1 << 16].tobytes())
        if skip == min(len(raw), 1 << 16) and len(raw) > (1 << 16):
            raise ValueError("read_bed: more than 64 KiB of header lines")
        raw = raw[skip:]
        if isinstance(config, str):
            config = Table(config)
        if fields is None:
            first = raw[: 4096].tobytes().split(b"\n", 1)[0]
            fields = max(3, min(6, first.count(b"\t") + 1))
        if not 3 <= fields <= 6:
            raise ValueError("fields must be 3..6")
        if len(raw) and not (raw.flags.writeable and raw.flags.c_contiguous):
            raw = raw.copy()                                             # torch wants a writable, contiguous source
        text = torch.from_numpy(raw).to(self.device) if len(raw) else self._empty(0, torch.uint8)
        n_lines = ctypes.c_int64()
        self._check(self.lib.giqlb_bed_count(self.ctx, _ptr(text), text.numel(), ctypes.byref(n_lines)))
        n = n_lines.value
        start, end, offs = self._empty(n, torch.int32), self._empty(n, torch.int32), self._empty(n + 1, torch.int32)
        name_offs = self._empty(n + 1, torch.int32) if fields >= 4 else None
        score = self._empty(n, torch.int32) if fields >= 5 else None
        strand = self._empty(n, torch.uint8) if fields >= 6 else None
        if n == 0:
            offs.zero_()
            t = DeviceTable(self, 0, config, start, end, strand, offs, self._empty(1, torch.uint8))
            t.bed_name, t.bed_score = None, score
            return t
        chrom_cap, name_cap = 8 * n, 16 * n if fields >= 4 else 0
        ct, nt = ctypes.c_int64(), ctypes.c_int64()
        for _ in range(2):
            cbytes = self._empty(chrom_cap + 16, torch.uint8)           # + 16: the dictionary kernels read whole 16-byte vectors
            nbytes = self._empty(name_cap + 16, torch.uint8) if fields >= 4 else None
            st = self.lib.giqlb_bed_parse(self.ctx, _ptr(text), text.numel(), n, fields, _ptr(offs), _ptr(cbytes), chrom_cap, _ptr(start), _ptr(end),
                                          _ptr(name_offs), _ptr(nbytes), name_cap, _ptr(score), _ptr(strand), ctypes.byref(ct), ctypes.byref(nt))
            if st != _cabi.ERR_CAPACITY:
                break
            chrom_cap, name_cap = ct.value, nt.value
        self._check(st)
        t = DeviceTable(self, n, config, start, end, strand, offs, cbytes[: max(ct.value, 1) + 16])
        t.bed_name = (name_offs, nbytes[: max(nt.value, 1)]) if fields >= 4 else None
        t.bed_score = score
        return t

    def table_from_buffers(self, chrom_offsets, chrom_bytes, start, end, strand=None,
                           config: Table | None = None) -> DeviceTable:
        """Table from columnar HOST buffers already in Arrow layout (utf8 offsets int32[n+1] + bytes, int32
        start/end, optional uint8 strand codes): numpy arrays or (pinned) torch tensors.  Only copies — no
        per-row Python work — so it is the ingest path for large inputs."""
        def dev(x, dt):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            if t.dtype != dt:
                raise ValueError(f"expected dtype {dt}, got {t.dtype}")
            return t.to(self.device, non_blocking=True)

        offs = dev(chrom_offsets, torch.int32)
        s, e = dev(start, torch.int32), dev(end, torch.int32)
        if offs.numel() != s.numel() + 1 or e.numel() != s.numel():
            raise ValueError("column lengths disagree")
        return DeviceTable(self, int(s.numel()), config, s, e, None if strand is None else dev(strand, torch.uint8),
                           offs, dev(chrom_bytes, torch.uint8))

    def table_from_codes(self, codes, start, end, names: Sequence[str] | Dictionary, strand=None,
                         config: Table | None = None) -> DeviceTable:
        """Table whose chromosome column is already dictionary-encoded (``names`` sorted bytewise)."""
        d = names if isinstance(names, Dictionary) else Dictionary(names)
        if list(d.names) != sorted(d.names, key=lambda s: s.encode()):
            raise ValueError("dictionary names must be sorted bytewise")

        c = codes.to(self.device, torch.int32).contiguous() if isinstance(codes, torch.Tensor) else self._dev(_as_i32(codes, "codes"))
        s = start.to(self.device, torch.int32).contiguous() if isinstance(start, torch.Tensor) else self._dev(_as_i32(start, "start"))
        e = end.to(self.device, torch.int32).contiguous() if isinstance(end, torch.Tensor) else self._dev(_as_i32(end, "end"))
        st = None
        if strand is not None:
            st = strand.to(self.device, torch.uint8).contiguous() if isinstance(strand, torch.Tensor) else self._dev(np.ascontiguousarray(strand, dtype=np.uint8))
        return DeviceTable(self, int(s.numel()), config, s, e, st, None, None, codes=c, dictionary=d)

    def encode(self, a: DeviceTable, b: Optional[DeviceTable] = None) -> Dictionary:
        """Dictionary-encode the chromosome column(s) against ONE shared dictionary (giqlb_dict_encode)."""
        if b is None or b is a:
            if a.dictionary is not None:
                return a.dictionary
        elif a.dictionary is not None and a.dictionary is b.dictionary:
            return a.dictionary
        tabs = [a] if (b is None or b is a) else [a, b]
        for t in tabs:
            if t.chrom_offsets is None:
                raise ValueError("tables encoded against different dictionaries cannot be combined; "
                                 "build them with the same Dictionary or from string columns")
        ca = self._empty(a.n, torch.int32)
        cb = self._empty(tabs[-1].n, torch.int32) if len(tabs) == 2 else None
        h = ctypes.c_void_p()
        zero = ctypes.c_void_p()
        t1 = tabs[-1] if len(tabs) == 2 else None
        self._check(self.lib.giqlb_dict_encode(
            self.ctx, _ptr(a.chrom_offsets), _ptr(a.chrom_bytes), a.n,
            _ptr(t1.chrom_offsets) if t1 else zero, _ptr(t1.chrom_bytes) if t1 else zero, t1.n if t1 else 0,
            _ptr(ca), _ptr(cb) if cb is not None else zero, ctypes.byref(h)))
        d = self._take_dictionary(h)
        a.codes, a.dictionary = ca, d
        if t1 is not None:
            t1.codes, t1.dictionary = cb, d
        return d

    # ---- INTERSECTS -------------------------------------------------------------------------------
    def _pair_indexes(self, a: DeviceTable, b: DeviceTable):
        d = self.encode(a, b)
        lay = _Layout(self, len(d))
        lay.observe(a)
        if b is not a:
            lay.observe(b)
        lay.finalize()
        ia = _Index(self, lay, a)
        ib = ia if b is a else _Index(self, lay, b)
        return lay, ia, ib

    def intersects(self, left: DeviceTable, right: DeviceTable, mode: str = "auto") -> PairResult:
        """Column-to-column INTERSECTS inner join -> (left_row, right_row) pairs.

        reference: src/giql/expanders/intersects.py:149-154 (predicate) evaluated per chromosome by
        src/giql/expanders/intersects_duckdb.py:1283-1299; bag semantics, unordered.  ``mode``: see
        :meth:`_join_indexes`.
        """
        if mode != "two_pass" and self._pair_capacity > 0:
            res = self._intersects_one_call(left, right)
            if res is not None:
                return res
        lay, il, ir = self._pair_indexes(left, right)
        try:
            return self._join_indexes(il, ir, mode)
        finally:
            il.free()
            ir.free()
            lay.free()

    def intersects_batches(self, batches, mode: str = "auto", depth: int = 2):
        """INTERSECTS joins over a stream of HOST batches with the PCIe transfers of neighbouring batches overlapped.

        ``batches`` yields ``(left, right)``; each side is a mapping with the pinned host buffers ``table_from_buffers``
        takes (``offs``, ``bytes``, ``start``, ``end``[, ``strand``]).  Yields ``(n_pairs, left_rows, right_rows)`` per
        batch, in order: pinned host int32 tensors from a ring of ``depth`` result buffers (a result stays valid until
        ``depth - 1`` further results have been taken).  Per batch the work is the same as ``table_from_buffers`` x2 + ``intersects`` + the
        copy of the pairs to the host; what changes is the schedule: batch i+1 is uploaded on its own stream while batch
        i is computed, and batch i's pairs travel back on a third stream while batch i+1 runs — PCIe is full duplex, so
        the step time drops from upload + compute + download to about max(upload, compute + download).
        """
        if depth < 2:
            raise ValueError("depth must be >= 2")
        main = self.stream
        # one upload and one download stream per engine, kept across calls: the caching allocator pools blocks per stream,
        # so fresh streams on every call would make every call pay cudaMalloc again (measured: 22 ms instead of 8.6 ms a batch)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        s_in, s_out = self._copy_streams

        def upload(batch):
            with torch.cuda.stream(s_in):
                tabs = []
                for side in batch:
                    tabs.append(self.table_from_buffers(side["offs"], side["bytes"], side["start"], side["end"], side.get("strand")))
                ev = torch.cuda.Event()
                ev.record(s_in)
            return tabs[0], tabs[1], ev

        ring = self.__dict__.setdefault("_result_ring", [])          # pinned (left_rows, right_rows) result buffers, kept across calls
        ring.extend([None] * (depth - len(ring)))
        # (slot, n, done event, tensors kept alive) not yet yielded.  Inputs live in s_in's allocator pool and pairs in the
        # engine stream's; holding them until the download event has completed makes their reuse safe without
        # record_stream (whose deferred frees made the caching allocator fall back to cudaMalloc / cudaFree every batch).
        inflight = []
        it = iter(batches)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            tl, tr, ev = nxt
            try:
                nxt = upload(next(it))                               # batch i+1 starts travelling before batch i is computed
            except StopIteration:
                nxt = None
            main.wait_event(ev)
            res = self.intersects(tl, tr, mode=mode)
            n = len(res)
            done_c = torch.cuda.Event()
            done_c.record(main)
            slot = i % depth
            if len(inflight) >= depth - 1:
                s0, n0, e0, keep = inflight.pop(0)                   # hand the oldest result out before a slot is reused
                e0.synchronize()
                del keep
                yield n0, ring[s0][0][:n0], ring[s0][1][:n0]
            if ring[slot] is None or ring[slot][0].numel() < n:
                cap = max(int(n * 1.25), 1024)
                ring[slot] = (torch.empty(cap, dtype=torch.int32, pin_memory=True), torch.empty(cap, dtype=torch.int32, pin_memory=True))
            s_out.wait_event(done_c)
            with torch.cuda.stream(s_out):
                ring[slot][0][:n].copy_(res.left_rows[:n], non_blocking=True)
                ring[slot][1][:n].copy_(res.right_rows[:n], non_blocking=True)
                done = torch.cuda.Event()
                done.record(s_out)
            inflight.append((slot, n, done, (tl, tr, res)))
            del tl, tr, res
            i += 1
        while inflight:
            s0, n0, e0, keep = inflight.pop(0)
            e0.synchronize()
            del keep
            yield n0, ring[s0][0][:n0], ring[s0][1][:n0]

    def _table_desc(self, t: DeviceTable) -> "_cabi.TableDesc":
        sadj, eadj = t.offsets
        if t.codes is None:
            t.codes = self._empty(t.n, torch.int32)
            valid = 0
        else:
            valid = 1 if t.dictionary is not None else 0
        d = _cabi.TableDesc()
        d.chrom_offsets = t.chrom_offsets.data_ptr() if t.chrom_offsets is not None else None
        d.chrom_bytes = t.chrom_bytes.data_ptr() if t.chrom_bytes is not None else None
        d.codes, d.codes_valid = t.codes.data_ptr(), valid
        d.start, d.end, d.n, d.start_adj, d.end_adj = t.start.data_ptr(), t.end.data_ptr(), t.n, sadj, eadj
        return d

    def _intersects_one_call(self, left: DeviceTable, right: DeviceTable) -> Optional[PairResult]:
        """The whole path through giqlb_intersects_tables (one FFI crossing).  Returns None when the tables cannot take
        it (one encoded, one not; encoded against different dictionaries) and the step-by-step path must run."""
        enc_l, enc_r = left.dictionary is not None, right.dictionary is not None
        if enc_l != enc_r or (enc_l and left.dictionary is not right.dictionary):
            return None
        if not enc_l and (left.chrom_offsets is None or right.chrom_offsets is None):
            return None
        dl = self._table_desc(left)
        dr = dl if right is left else self._table_desc(right)
        n_codes = len(left.dictionary) if enc_l else 0
        cap = max(self._pair_capacity, 1024)
        total, dict_h = ctypes.c_int64(), ctypes.c_void_p()
        for _ in range(2):
            out_l, out_r = self._empty(cap, torch.int32), self._empty(cap, torch.int32)
            st = self.lib.giqlb_intersects_tables(self.ctx, ctypes.byref(dl), ctypes.byref(dr), n_codes, _ptr(out_l), _ptr(out_r), cap,
                                                  ctypes.byref(total), ctypes.byref(dict_h))
            if dict_h.value:
                d = self._take_dictionary(dict_h)
                if st in (_cabi.OK, _cabi.ERR_CAPACITY):      # the C side marks the codes valid on these two only (pipeline.cu):
                    left.dictionary = right.dictionary = d    # on any other status the code columns may still hold parked slots
                    n_codes, enc_l = len(d), True
            if st == _cabi.OK:
                return PairResult(out_l[: total.value], out_r[: total.value])
            if st != _cabi.ERR_CAPACITY:
                self._check(st)
            cap = total.value
            self._pair_capacity = int(cap * 1.25) + 1024
        raise EngineError(_cabi.ERR_INTERNAL, "intersects_tables: capacity retry did not converge")

    def intersects_shard(self, left: DeviceTable, right: DeviceTable, world: int, rank: int) -> PairResult:
        """Rank ``rank``'s share of the join under genome-range sharding, through giqlb_intersects_tables_sharded (one FFI
        crossing: encode, layout, device-side cuts, fused select + keygen, index, single-pass join)."""
        enc_l, enc_r = left.dictionary is not None, right.dictionary is not None
        if enc_l != enc_r or (enc_l and left.dictionary is not right.dictionary) or right is left:
            raise ValueError("sharded join needs two distinct tables, both un-encoded or encoded against one dictionary")
        dl, dr = self._table_desc(left), self._table_desc(right)
        n_codes = len(left.dictionary) if enc_l else 0
        cap = max(self._pair_capacity, 1024)
        total, dict_h = ctypes.c_int64(), ctypes.c_void_p()
        for _ in range(2):
            out_l, out_r = self._empty(cap, torch.int32), self._empty(cap, torch.int32)
            st = self.lib.giqlb_intersects_tables_sharded(self.ctx, ctypes.byref(dl), ctypes.byref(dr), n_codes, world, rank,
                                                          _ptr(out_l), _ptr(out_r), cap, ctypes.byref(total), ctypes.byref(dict_h))
            if dict_h.value:
                d = self._take_dictionary(dict_h)
                if st in (_cabi.OK, _cabi.ERR_CAPACITY):      # see _intersects_one_call
                    left.dictionary = right.dictionary = d
                    n_codes = len(d)
            if st == _cabi.OK:
                return PairResult(out_l[: total.value], out_r[: total.value])
            if st != _cabi.ERR_CAPACITY:
                self._check(st)
            cap = total.value
            self._pair_capacity = int(cap * 1.25) + 1024
        raise EngineError(_cabi.ERR_INTERNAL, "intersects_tables_sharded: capacity retry did not converge")

    def _take_dictionary(self, h: ctypes.c_void_p) -> Dictionary:
        try:
            names = []
            for code in range(self.lib.giqlb_dict_size(h)):
                p, ln = ctypes.c_void_p(), ctypes.c_int32()
                self.lib.giqlb_dict_get(h, code, ctypes.byref(p), ctypes.byref(ln))
                names.append(ctypes.string_at(p, ln.value).decode("utf-8", "surrogateescape"))
        finally:
            self.lib.giqlb_dict_free(h)
        return Dictionary(names)

    def _join_indexes(self, il: _Index, ir: _Index, mode: str = "auto") -> PairResult:
        """INTERSECTS over two indexes built on one layout.

        ``two_pass``: giqlb_intersects_plan (count + scan, exact size) then giqlb_intersects_emit.
        ``single_pass``: giqlb_intersects_join into a buffer sized from the engine's running capacity estimate — one
        fused kernel counts, offsets and emits; when the estimate is too small the library reports the exact size
        and the join is repeated once into an exact buffer.  ``auto`` = two_pass until an estimate exists.
        """
        if mode not in ("auto", "two_pass", "single_pass"):
            raise ValueError(f"unknown join mode {mode!r}")
        if mode == "two_pass" or (mode == "auto" and self._pair_capacity == 0):
            plan, total = ctypes.c_void_p(), ctypes.c_int64()
            self._check(self.lib.giqlb_intersects_plan(self.ctx, il.h, ir.h, ctypes.byref(plan), ctypes.byref(total)))
            try:
                out_l = self._empty(total.value, torch.int32)
                out_r = self._empty(total.value, torch.int32)
                self._check(self.lib.giqlb_intersects_emit(self.ctx, plan, _ptr(out_l), _ptr(out_r)))
            finally:
                self.lib.giqlb_joinplan_free(self.ctx, plan)
            self._pair_capacity = max(self._pair_capacity, int(total.value * 1.25) + 1024)
            return PairResult(out_l, out_r)
        cap = max(self._pair_capacity, 1024)
        total = ctypes.c_int64()
        for _ in range(2):
            out_l = self._empty(cap, torch.int32)
            out_r = self._empty(cap, torch.int32)
            st = self.lib.giqlb_intersects_join(self.ctx, il.h, ir.h, _ptr(out_l), _ptr(out_r), cap, ctypes.byref(total))
            if st == _cabi.OK:
                return PairResult(out_l[: total.value], out_r[: total.value])
            if st != _cabi.ERR_CAPACITY:
                self._check(st)
            cap = total.value                      # exact size reported by the library
            self._pair_capacity = int(cap * 1.25) + 1024
        raise EngineError(_cabi.ERR_INTERNAL, "intersects_join: capacity retry did not converge")

    # ---- genome-range sharding (orchestration in giql_b200/sharding.py) -------------
    def shard_cuts(self, lay: _Layout, left: DeviceTable, world: int) -> tuple[list[int], int]:
        """Linear-coordinate cuts balancing the LEFT interval count over ``world`` ranges, and the widest
        left interval (giqlb_shard_cuts).  Deterministic, so every rank computes the same cuts."""
        cuts = (ctypes.c_uint32 * (world + 1))()
        mw = ctypes.c_uint32()
        sadj, eadj = left.offsets
        self._check(self.lib.giqlb_shard_cuts(self.ctx, lay.h, _ptr(left.codes), _ptr(left.start), _ptr(left.end), left.n,
                                              sadj, eadj, world, cuts, ctypes.byref(mw)))
        return list(cuts), int(mw.value)

    def shard_select(self, lay: _Layout, t: DeviceTable, gs_lo: int, gs_hi: int, ge_gt: int) -> torch.Tensor:
        """Ascending ids of the rows with gs_lo <= linear start < gs_hi and linear end > ge_gt."""
        rows = self._empty(t.n, torch.int32)
        n_out = ctypes.c_int64()
        sadj, eadj = t.offsets
        self._check(self.lib.giqlb_shard_select(self.ctx, lay.h, _ptr(t.codes), _ptr(t.start), _ptr(t.end), t.n, sadj, eadj,
                                                gs_lo, gs_hi, ge_gt, _ptr(rows), ctypes.byref(n_out)))
        return rows[: n_out.value]

    def _spatial(self, left: DeviceTable, right: DeviceTable, op: int) -> PairResult:
        lay, il, ir = self._pair_indexes(left, right)
        try:
            total = ctypes.c_int64()
            st = self.lib.giqlb_spatial_join(self.ctx, il.h, ir.h, op, None, None, 0, ctypes.byref(total))   # sizing call
            if st not in (_cabi.OK, _cabi.ERR_CAPACITY):
                self._check(st)
            out_l, out_r = self._empty(total.value, torch.int32), self._empty(total.value, torch.int32)
            if total.value:
                self._check(self.lib.giqlb_spatial_join(self.ctx, il.h, ir.h, op, _ptr(out_l), _ptr(out_r), total.value, ctypes.byref(total)))
            return PairResult(out_l, out_r)
        finally:
            il.free()
            ir.free()
            lay.free()

    _CMP_OPS = {"=": 0, "==": 0, "!=": 1, "<>": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}
    _CMP_DTYPES = {torch.uint8: 1, torch.int32: 2, torch.int64: 3, torch.float32: 4, torch.float64: 5}

    def filter_pairs(self, pairs: PairResult, left_col, right_col, op: str, *, left_valid=None, right_valid=None,
                     strand: bool = False) -> PairResult:
        """Residual predicate over join pairs: keeps ``left_col[l] <op> right_col[r]`` (both non-NULL), order preserved.

        The extra ON-conjunct of intersects_duckdb.py:1239-1243 (``AND a.score > b.score``).  Columns: device or host
        arrays of one common dtype (uint8 / int32 / int64 / float32 / float64), indexed by table row; ``*_valid``:
        optional 0/1 masks for NULLs; ``strand=True`` treats uint8 columns as strand codes (255 = NULL)."""
        if op not in self._CMP_OPS:
            raise ValueError(f"unknown comparison {op!r}")

        def dev(x, dt=None):
            if x is None:
                return None
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            if dt is not None:
                t = t.to(dt)
            return t.to(self.device).contiguous()

        lc, rc = dev(left_col), dev(right_col)
        if lc.dtype != rc.dtype:
            common = torch.promote_types(lc.dtype, rc.dtype)
            lc, rc = lc.to(common), rc.to(common)
        if lc.dtype in (torch.int8, torch.int16, torch.bool):
            lc, rc = lc.to(torch.int32), rc.to(torch.int32)
        if lc.dtype not in self._CMP_DTYPES:
            raise ValueError(f"unsupported column dtype {lc.dtype}")
        dtype = 0 if strand else self._CMP_DTYPES[lc.dtype]
        if strand and lc.dtype != torch.uint8:
            raise ValueError("strand columns must be uint8 codes")
        lv, rv = dev(left_valid, torch.uint8), dev(right_valid, torch.uint8)
        n = len(pairs)
        out_l, out_r = self._empty(n, torch.int32), self._empty(n, torch.int32)
        kept = ctypes.c_int64()
        self._check(self.lib.giqlb_pairs_filter(self.ctx, _ptr(pairs.left_rows), _ptr(pairs.right_rows), n, _ptr(lc), _ptr(rc), dtype,
                                                None if lv is None else _ptr(lv), None if rv is None else _ptr(rv), self._CMP_OPS[op],
                                                _ptr(out_l), _ptr(out_r), ctypes.byref(kept)))
        return PairResult(out_l[: kept.value], out_r[: kept.value])

    def intersects_stranded(self, left: DeviceTable, right: DeviceTable, same: bool = True, mode: str = "auto") -> PairResult:
        """INTERSECTS join restricted to pairs on the same (``a.strand = b.strand``, bedtools -s) or on opposite
        (``a.strand != b.strand``, bedtools -S) strands — docs/recipes/bedtools-migration.rst:271-307.  SQL semantics:
        plain (in)equality of the strand literals, NULL strands never qualify."""
        if left.strand is None or right.strand is None:
            raise ValueError("both tables need a strand column")
        pairs = self.intersects(left, right, mode=mode)
        return self.filter_pairs(pairs, left.strand, right.strand, "=" if same else "!=", strand=True)

    def contains(self, left: DeviceTable, right: DeviceTable) -> PairResult:
        """``left.interval CONTAINS right.interval`` join: l.start <= r.start AND l.end >= r.end on one chromosome
        (reference: src/giql/expanders/intersects.py:155-160)."""
        return self._spatial(left, right, _cabi.OP_CONTAINS)

    def within(self, left: DeviceTable, right: DeviceTable) -> PairResult:
        """``left.interval WITHIN right.interval`` join: l.start >= r.start AND l.end <= r.end
        (reference: src/giql/expanders/intersects.py:161-166)."""
        return self._spatial(left, right, _cabi.OP_WITHIN)

    def _probe(self, fn, left: DeviceTable, right: DeviceTable, out: torch.Tensor):
        d = self.encode(left, right)
        lay = _Layout(self, len(d))
        lay.observe(right)
        lay.finalize()
        ir = _Index(self, lay, right)
        try:
            sadj, eadj = left.offsets
            self._check(fn(self.ctx, lay.h, _ptr(left.codes), _ptr(left.start), _ptr(left.end), left.n, sadj, eadj, ir.h, _ptr(out)))
        finally:
            ir.free()
            lay.free()
        return out

    def exists(self, left: DeviceTable, right: DeviceTable) -> torch.Tensor:
        """uint8[n_left]: 1 iff the left row overlaps at least one right row (the EXISTS of
        intersects_duckdb.py:1254-1282)."""
        return self._probe(self.lib.giqlb_intersects_exists, left, right, self._empty(left.n, torch.uint8))

    def _select(self, flags: torch.Tensor, want: int) -> torch.Tensor:
        out = self._empty(flags.numel(), torch.int32)
        n_out = ctypes.c_int64()
        self._check(self.lib.giqlb_select_rows(self.ctx, _ptr(flags), flags.numel(), want, _ptr(out), ctypes.byref(n_out)))
        return out[: n_out.value]

    def semi(self, left: DeviceTable, right: DeviceTable) -> torch.Tensor:
        """SEMI JOIN: ids of left rows with >= 1 match, one per qualifying ROW, ascending
        (intersects_duckdb.py:1254-1282, :1259-1262 no dedup)."""
        return self._select(self.exists(left, right), 1)

    def anti(self, left: DeviceTable, right: DeviceTable) -> torch.Tensor:
        """ANTI JOIN: ids of left rows with no match; rows on chromosomes absent from ``right`` are kept
        (partition rule intersects_duckdb.py:1317-1330, tests/test_duckdb_iejoin.py:5873-5919)."""
        return self._select(self.exists(left, right), 0)

    def count_overlaps(self, left: DeviceTable, right: DeviceTable) -> torch.Tensor:
        """int64[n_left]: number of right rows overlapping each left row (count_overlaps fast path,
        intersects_duckdb.py:432-548; the GROUP BY over distinct left keys is a host-side sum of these)."""
        return self._probe(self.lib.giqlb_intersects_count_rows, left, right, self._empty(left.n, torch.int64))

    # ---- NEAREST -------------------------------------------------------------------------------------
    def nearest(self, query: DeviceTable, target: DeviceTable, k: int = 1, max_distance: Optional[int] = None,
                stranded: bool = False, signed: bool = False) -> "NearestResult":
        """``query CROSS JOIN LATERAL NEAREST(target, reference := query.interval, k, max_distance, stranded, signed)``
        -> (query_row, target_row, distance) rows, at most k per query row.
        reference: src/giql/expanders/nearest.py:255-397; parameters :240-252 (k defaults to 1, must be >= 1)."""
        k = int(k)
        if k < 1:
            raise ValueError("k must be a positive integer")
        if max_distance is not None and int(max_distance) < 0:
            raise ValueError("max_distance must be non-negative")
        # the strand filter and the stranded CASE apply only when BOTH tables carry a strand column (nearest.py:304-307)
        use_strand = bool(stranded) and query.strand is not None and target.strand is not None
        self.encode(query, target)
        cap = k * query.n
        out_q, out_t, out_d = (self._empty(cap, torch.int32) for _ in range(3))
        out_v = self._empty(cap, torch.uint8)
        n_out = ctypes.c_int64(0)
        if query.n and target.n:
            lay = _Layout(self, len(query.dictionary))
            lay.observe(target)
            lay.finalize()
            ix = _Index(self, lay, target, with_strand=use_strand, tie_by_end=True)
            try:
                if use_strand:
                    self._check(self.lib.giqlb_index_group_by_strand(self.ctx, ix.h))
                qs, qe = query.offsets
                self._check(self.lib.giqlb_nearest(
                    self.ctx, lay.h, _ptr(query.codes), _ptr(query.start), _ptr(query.end),
                    _ptr(query.strand) if use_strand else None, query.n, qs, qe, ix.h, k,
                    -1 if max_distance is None else int(max_distance), int(use_strand), int(bool(signed)),
                    _ptr(out_q), _ptr(out_t), _ptr(out_d), _ptr(out_v), cap, ctypes.byref(n_out)))
            finally:
                ix.free()
                lay.free()
        m = n_out.value
        return NearestResult(out_q[:m], out_t[:m], out_d[:m], out_v[:m])

    # ---- CLUSTER / MERGE ---------------------------------------------------------------------------
    def _self_index(self, t: DeviceTable, stranded: bool, raw: bool, tie_by_end: bool = False, row_ids: bool = True):
        d = self.encode(t)
        lay = _Layout(self, len(d))
        lay.observe(t, raw=raw)
        lay.finalize()
        ix = _Index(self, lay, t, raw=raw, with_strand=stranded, tie_by_end=tie_by_end, row_ids=row_ids)
        if stranded:
            self._check(self.lib.giqlb_index_group_by_strand(self.ctx, ix.h))
        return lay, ix

    def strand_literal(self, code: int):
        """Inverse of :meth:`strand_codes` for one code (None for SQL NULL)."""
        if code == _cabi.STRAND_NULL:
            return None
        for lit, c in self._strand_codes.items():
            if c == code:
                return lit
        raise KeyError(code)

    @staticmethod
    def predicate_codes(cur_values, prev_values):
        """Order-preserving int32 codes of the two operand columns of ``predicate := cur <op> PREV(prev)``, ranked together
        so that codes compare like the values; None / NaN-free NULLs become GIQLB_KEY_NULL."""
        cur, prev = list(cur_values), list(prev_values)
        try:
            values = sorted({v for v in cur if v is not None} | {v for v in prev if v is not None})
        except TypeError as exc:
            raise ValueError(f"predicate operands are not comparable: {exc}") from None
        rank = {v: i for i, v in enumerate(values)}
        null = _cabi.KEY_NULL
        enc = lambda col: np.fromiter((null if v is None else rank[v] for v in col), dtype=np.int32, count=len(col))
        return enc(cur), enc(prev)

    def _predicate_codes_device(self, cur_values, prev_values):
        """The same codes for NUMERIC operand columns (numpy arrays or torch tensors of one integer or one floating kind),
        ranked on the device: no per-row Python, usable at the 200 M rows of BASELINE configs[4].  NaN is the NULL of a
        floating column.  Returns None for anything else (lists, objects, mixed kinds): the host path above ranks those."""
        cols = []
        for v in (cur_values, prev_values):
            if isinstance(v, np.ndarray) and v.dtype.kind in "iuf" and v.ndim == 1:
                v = torch.from_numpy(np.ascontiguousarray(v))
            if not isinstance(v, torch.Tensor) or v.ndim != 1 or v.dtype == torch.bool or v.dtype.is_complex:
                return None
            cols.append(v.to(self.device, non_blocking=True))
        if cols[0].dtype.is_floating_point != cols[1].dtype.is_floating_point:
            return None
        both = torch.cat([c.to(torch.float64 if c.dtype.is_floating_point else torch.int64) for c in cols])
        codes = torch.full((both.numel(),), _cabi.KEY_NULL, dtype=torch.int32, device=self.device)
        known = ~torch.isnan(both) if both.dtype.is_floating_point else None
        vals = both[known] if known is not None else both
        inv = torch.unique(vals, sorted=True, return_inverse=True)[1].to(torch.int32)
        if known is not None:
            codes[known] = inv
        else:
            codes = inv
        n = cols[0].numel()
        return codes[:n].contiguous(), codes[n:].contiguous()

    def cluster(self, t: DeviceTable, distance: int = 0, stranded: bool = False, predicate=None) -> torch.Tensor:
        """CLUSTER(interval[, distance][, stranded := true][, predicate := cur <op> PREV(prev)]) -> int32 cluster id per
        input row (1-based, restarting per chrom[, strand] partition).  reference: src/giql/expanders/cluster.py:205-310,
        430-434 (raw table coordinates: expressions.py:223-226).  ``predicate``: ``(cur_values, op, prev_values)`` with one
        value per input row (host sequences) — the row also has to satisfy ``cur <op> PREV(prev)`` against its immediate
        sorted predecessor to stay in the cluster (cluster.py:281-299)."""
        out = self._empty(t.n, torch.int32)
        if t.n == 0:
            return out
        keys = None
        if predicate is not None:
            cur_values, op, prev_values = predicate
            if op not in self._CMP_OPS:
                raise ValueError(f"unknown comparison {op!r}")
            if len(cur_values) != t.n or len(prev_values) != t.n:
                raise ValueError("predicate columns must have one value per row")
            dev_keys = self._predicate_codes_device(cur_values, prev_values)
            if dev_keys is not None:
                keys = (dev_keys[0], dev_keys[1], self._CMP_OPS[op])
            else:
                kc, kp = self.predicate_codes(cur_values, prev_values)
                keys = (self._dev(kc), self._dev(kp), self._CMP_OPS[op])
        lay, ix = self._self_index(t, stranded, raw=True)
        try:
            if keys is None and not stranded:
                # plain index over the whole table: few clusters are resolved per input row by search (giqlb_cluster_table)
                self._check(self.lib.giqlb_cluster_table(self.ctx, ix.h, int(distance or 0), _ptr(t.codes), _ptr(t.start), t.n, 0, _ptr(out)))
            elif keys is None:
                self._check(self.lib.giqlb_cluster(self.ctx, ix.h, int(distance or 0), _ptr(out)))
            else:
                self._check(self.lib.giqlb_cluster_predicate(self.ctx, ix.h, int(distance or 0), _ptr(keys[0]), _ptr(keys[1]), keys[2], _ptr(out)))
        finally:
            ix.free()
            lay.free()
        return out

    def merge(self, t: DeviceTable, distance: int = 0, stranded: bool = False) -> "MergeResult":
        """MERGE(interval[, distance][, stranded := true]) -> merged regions with COUNT(*), ordered by chrom, start
        (strand block first when stranded).  reference: src/giql/expanders/merge.py:183-439."""
        d = self.encode(t)
        n = t.n
        code, start, end = (self._empty(n, torch.int32) for _ in range(3))
        strand = self._empty(n, torch.uint8) if stranded else None
        count = self._empty(n, torch.int64)
        m = ctypes.c_int64(0)
        if n:
            lay, ix = self._self_index(t, stranded, raw=True, row_ids=False)   # regions, not rows: no row-id payload in the sort
            try:
                self._check(self.lib.giqlb_merge(self.ctx, ix.h, int(distance or 0), _ptr(code), _ptr(start), _ptr(end),
                                                 _ptr(strand), _ptr(count), n, ctypes.byref(m)))
            finally:
                ix.free()
                lay.free()
        k = m.value
        return MergeResult(self, d, code[:k], start[:k], end[:k], None if strand is None else strand[:k], count[:k])

    # ---- DISJOIN -------------------------------------------------------------------------------------
    def disjoin(self, target: DeviceTable, reference: Optional[DeviceTable] = None) -> "DisjoinResult":
        """DISJOIN(target[, reference]) -> (target_row, segment_start, segment_end) in canonical coordinates.
        reference: src/giql/expanders/disjoin.py:102-202 (self-reference skips the coverage filter, :184-190)."""
        ref = target if reference is None else reference
        self.encode(target, ref)
        empty = lambda: DisjoinResult(*(self._empty(0, torch.int32) for _ in range(3)))
        if target.n == 0:
            return empty()
        lay = _Layout(self, len(target.dictionary))
        lay.observe(ref)
        lay.finalize()
        ix = _Index(self, lay, ref)
        try:
            ts, te = target.offsets
            args = (self.ctx, lay.h, _ptr(target.codes), _ptr(target.start), _ptr(target.end), target.n, ts, te, ix.h,
                    int(reference is None))
            n_out = ctypes.c_int64(0)
            st = self.lib.giqlb_disjoin(*args, None, None, None, 0, ctypes.byref(n_out))   # sizing call
            if st not in (_cabi.OK, _cabi.ERR_CAPACITY):
                self._check(st)
            m = n_out.value
            if m == 0:
                return empty()
            row, s, e = (self._empty(m, torch.int32) for _ in range(3))
            self._check(self.lib.giqlb_disjoin(*args, _ptr(row), _ptr(s), _ptr(e), m, ctypes.byref(n_out)))
            return DisjoinResult(row, s, e)
        finally:
            ix.free()
            lay.free()

    # ---- DISTANCE ---------------------------------------------------------------------------------
    def distance(self, a: DeviceTable, b: DeviceTable, a_rows: Optional[torch.Tensor] = None,
                 b_rows: Optional[torch.Tensor] = None, stranded: bool = False, signed: bool = False):
        """DISTANCE(a.interval, b.interval) for paired rows -> (int32 distance, uint8 valid[0 = SQL NULL]).

        ``a_rows`` / ``b_rows`` select the paired rows (e.g. a join result); omitted = row i with row i.
        reference: src/giql/expanders/_distance.py:67-117.  The stranded CASE is used only when both
        tables carry a strand column (_distance.py:67).
        """
        self.encode(a, b)

        def rows32(rows, limit):
            if rows is None:
                return None, limit
            r = rows if rows.dtype == torch.int32 else rows.to(torch.int32)
            return r.contiguous(), int(r.numel())

        ra, na = rows32(a_rows, a.n)
        rb, nb = rows32(b_rows, b.n)
        if na != nb:
            raise ValueError("paired row selections differ in length")
        n = na
        use_strand = bool(stranded) and a.strand is not None and b.strand is not None
        out = self._empty(n, torch.int32)
        valid = self._empty(n, torch.uint8)
        (asj, aej), (bsj, bej) = a.offsets, b.offsets
        # the row selections are gathered inside the kernel (giqlb_distance_rows): no gathered copies of the six columns
        self._check(self.lib.giqlb_distance_rows(
            self.ctx, _ptr(a.codes), _ptr(a.start), _ptr(a.end), _ptr(a.strand) if use_strand else None, a.n,
            _ptr(b.codes), _ptr(b.start), _ptr(b.end), _ptr(b.strand) if use_strand else None, b.n,
            _ptr(ra) if ra is not None else None, _ptr(rb) if rb is not None else None,
            n, int(use_strand), int(bool(signed)), asj, aej, bsj, bej, _ptr(out), _ptr(valid)))
        return out, valid
