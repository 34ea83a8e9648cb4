#!/usr/bin/env python
"""bench.py — the north-star measurement: overlap pairs/sec of the 10M x 10M INTERSECTS inner join.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--rows ROWS]

One *step* is one pass of the hot path over the synthetic hg38-shaped BED tables of BASELINE.json
configs[1] (`peaks(1e7, 20260103)` x `peaks(1e7, 20260104)`, row order shuffled):

    chromosome dictionary encoding  ->  genome layout  ->  LSD radix-sort index of both tables
    ->  join count  ->  scan  ->  join emit                      (all kernels of libgiqlb200.so)

`value`  — pairs / device time of the step, inputs (utf8 chrom + int32 start/end) resident in HBM.
`e2e`    — the same metric through the public host API (`Engine.table_from_buffers` + `Engine.intersects`
           + pairs copied back) with PINNED HOST buffers: H2D of the inputs and D2H of the pairs are inside
           the timed region, wall clock bracketed by synchronisation.
`roofline` — the dominant kernel group of the step, timed live with CUDA events on the engine's stream.
`cpu_baseline` — the C/OpenMP restatement in oracle/ on the host cores (NOT DuckDB: duckdb and sqlglot
           are absent from the image), rank 0, N=1 only.

`--impl reference` times that same CPU restatement as the reference arm (oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "overlap pairs/sec, 10Mx10M INTERSECTS inner join"
UNIT = "pairs/s"
SEED_L, SEED_R = 20260103, 20260104
ROWS_DEFAULT = 10_000_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- workload -------------------------------------------------------------------------------------------
def utf8_columns(codes: np.ndarray, names):
    """Arrow utf8 layout (int32 offsets[n+1], uint8 bytes) of names[codes]."""
    import pyarrow as pa

    arr = pa.DictionaryArray.from_arrays(pa.array(codes, type=pa.int32()), pa.array(list(names), type=pa.string())).cast(pa.string())
    bufs = arr.buffers()
    offs = np.frombuffer(bufs[1], dtype=np.int32)[: len(codes) + 1].copy()
    data = np.frombuffer(bufs[2], dtype=np.uint8)[: int(offs[-1])].copy()
    return offs, data


def make_workload(rows: int):
    from giql_b200 import synth

    t0 = time.time()
    A, B = synth.peaks(rows, SEED_L), synth.peaks(rows, SEED_R)
    for T in (A, B):
        T["offs"], T["bytes"] = utf8_columns(T["code"], T["names"])
    log(f"[bench] generated 2 x {rows} intervals in {time.time() - t0:.1f}s")
    return A, B


def workload_name(rows):
    return (f"C2 INTERSECTS inner join {rows}x{rows}, synthetic hg38 BED (24 chroms, log-normal widths median 500bp, "
            f"80% clustered starts, shuffled rows), seeds {SEED_L}/{SEED_R}")


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = float(os.environ.get("GIQLB_BENCH_CLOCK_PERIOD_S", "0.002"))):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # pragma: no cover
            log(f"[bench] NVML unavailable ({exc}); clocks not sampled")
            self.nv = None

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
              0x10: "sync_boost"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---- CPU restatement (cpu_baseline and the reference arm) -------------------------------------------------------
def cpu_join(A, B, threads: int):
    """One whole CPU join (bucket + per-chromosome sort + count + scan + emit) of the oracle port."""
    from oracle import c_oracle

    cap = getattr(cpu_join, "_cap", 0)
    t0 = time.perf_counter()
    total, _, _ = c_oracle.intersects_join_timed(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"],
                                                 len(A["names"]), cap, threads)
    if total > cap:   # first call sizes the output (untimed), second call is the measured one
        cpu_join._cap = int(total)
        t0 = time.perf_counter()
        total, _, _ = c_oracle.intersects_join_timed(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"],
                                                     len(A["names"]), int(total), threads)
    return int(total), time.perf_counter() - t0


def duckdb_baseline(A, B, threads: int, repeats: int = 3):
    """The reference's own path — giql.transpile(dialect="duckdb") executed by DuckDB on the host cores — when duckdb (and
    giql) are importable on this box; None otherwise (today: neither is installed in the image)."""
    from oracle import duckdb_oracle as do

    if not do.available():
        return None
    try:
        pairs, seconds, _ = do.intersects_join_timed(A, B, threads, plan="iejoin", repeats=repeats)
    except Exception as exc:   # noqa: BLE001 — a broken DuckDB install must not take the bench line down
        log(f"[bench] duckdb baseline failed ({type(exc).__name__}: {exc}); falling back to the C restatement")
        return None
    return {"value": pairs / seconds, "unit": UNIT, "cores": threads, "kind": "duckdb" if do.have_giql() else "duckdb-restated-sql",
            "sample": f"full workload, best of {repeats} after 1 warm-up ({seconds:.2f}s): {do.label(threads)}, "
                      "CREATE TEMP TABLE out AS <final SELECT of the per-chromosome IEJoin plan>", "pairs": pairs, "seconds": seconds}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.build()
    threads = c_oracle.default_threads()
    A, B = make_workload(args.rows)
    duck = duckdb_baseline(A, B, threads, repeats=max(1, min(args.steps, 3)))
    if duck is not None:
        # the true reference path exists on this box: GIQL's transpiled SQL on DuckDB
        line = {
            "impl": "reference", "metric": METRIC, "value": duck["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * duck["seconds"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.rows), "pairs": duck["pairs"], "note": "reference arm = " + duck["sample"]},
            "cpu_baseline": {k: duck[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": duck["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    for _ in range(args.warmup):
        cpu_join(A, B, threads)
    times, total = [], 0
    for _ in range(args.steps):
        total, dt = cpu_join(A, B, threads)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = total / (ms / 1e3)
    sample = f"full workload per step ({args.rows}x{args.rows} rows, {total} pairs)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args.rows), "pairs": total,
                   "note": "reference arm = C/OpenMP restatement of the SQL GIQL emits (oracle/c/giql_oracle.c); "
                           "duckdb and sqlglot are not installed in this image, so this is NOT DuckDB"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- B200 arm --------------------------------------------------------------------------------------------------
def algorithmic_bytes(phase: str, n_full: int, n_local: int, pairs: int, sort_passes: int):
    """Per-phase algorithmic bytes of ONE step on ONE rank.  n_full = rows of both tables (phases that
    read the whole replicated input), n_local = rows this rank indexes and joins (= n_full on one GPU)."""
    return {
        "dict": 8 * n_full,                        # read the utf8 name (~4.6 B) + offset, write an int32 code: counted as 8 B/row
        "layout": 12 * n_full,                     # read code/start/end
        "shard": 12 * n_full + 12 * n_local,       # sampled histogram + fused select/keygen: read the columns, write keys + row ids
        "keygen": 12 * n_local + 8 * n_local,      # read code/start/end, write linear start + linear end
        "sort_hist": 4 * n_local,                  # read keys once
        "sort_pass": int(16 * n_local * sort_passes),   # keys + row ids in and out, per 8-bit digit pass
        # after a full sort: read row id + gathered end, write end.  After the short-window sort (three passes): the finishing
        # pass reads key + row id + gathered end and writes key, row id and end
        "gather": (12 if sort_passes >= 4 else 24) * n_local,
        "join_count": 8 * n_local,                 # merge-path partition (or the count pass of the two-pass API)
        "join_scan": 0,
        "join_emit": 12 * n_local + 8 * pairs,
    }.get(phase, 0)



# ---- the other BASELINE configs (C3 / C4 / C5), reported beside the headline in the same JSON line ---------------------
def other_configs(eng, rank, world, hbm_peak, rows_scale=1.0):
    """BASELINE.json configs[2..4] on this run's GPUs: rows/s, algorithmic bytes and fraction of the measured
    HBM peak.  Inputs are generated on the device (giql_b200/synth_torch.py: the §8(d) distributions, a torch random stream)
    and arrive as dictionary codes; each timing is the median of 5 calls after 2 warm-ups (CUDA events on the engine's
    stream) and includes the index build of the target / own table.  N = 1: SEMI / ANTI 50M x 1M, NEAREST k = 1 / 5 stranded
    10M x 2M with DISTANCE, CLUSTER / MERGE over 200M reads.  N > 1: CLUSTER / MERGE over the same 200M reads, genome-range
    sharded over the N ranks (three small NCCL all-gathers, no row exchange), totals checked against the single-GPU answer."""
    import torch
    import torch.distributed as dist

    from giql_b200 import synth_torch as st

    out = []
    dev = eng.device

    def table(T, strand=False, dictionary=None):
        return eng.table_from_codes(T["code"], T["start"], T["end"], dictionary or T["names"], strand=T["strand_code"] if strand else None)

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            res = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(eng.stream)
            res = fn()
            b.record(eng.stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), res

    def emit(name, rows, ms, alg_bytes, **extra):
        gbs = alg_bytes / (ms * 1e6)
        line = {"config": name, "rows": rows, "ms": round(ms, 3), "rows_per_s": rows / (ms / 1e3), "algorithmic_bytes": int(alg_bytes),
                "gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)}
        line.update(extra)
        out.append(line)
        log(f"[bench] {name}: {ms:.3f} ms, {rows / (ms / 1e3) / 1e9:.2f} G rows/s, {gbs:.0f} GB/s = {gbs / hbm_peak:.3f} of HBM")

    n5 = int(200_000_000 * rows_scale)
    if world == 1:
        nq, nt = int(50_000_000 * rows_scale), int(1_000_000 * rows_scale)
        tq = table(st.peaks(nq, 20260105, dev))
        tt = table(st.peaks(nt, 20260106, dev), dictionary=tq.dictionary)
        ms, kept = timed(lambda: eng.semi(tq, tt))
        emit(f"C3 SEMI {nq} x {nt}", nq, ms, 12 * (nq + nt) + 4 * int(kept.numel()), kept=int(kept.numel()))
        ms, kept_a = timed(lambda: eng.anti(tq, tt))
        emit(f"C3 ANTI {nq} x {nt}", nq, ms, 12 * (nq + nt) + 4 * int(kept_a.numel()), kept=int(kept_a.numel()))
        assert int(kept.numel()) + int(kept_a.numel()) == nq
        del tq, tt, kept, kept_a
        nq, nt = int(10_000_000 * rows_scale), int(2_000_000 * rows_scale)
        tq = table(st.peaks(nq, 20260107, dev), strand=True)
        tt = table(st.genes(nt, 20260108, dev), strand=True, dictionary=tq.dictionary)
        for k in (1, 5):
            ms, res = timed(lambda: eng.nearest(tq, tt, k=k, stranded=True, signed=True))
            emit(f"C4 NEAREST k={k} stranded+signed {nq} x {nt} (distance emitted)", nq, ms, 13 * (nq + nt) + 12 * len(res), emitted=len(res))
        ms, _ = timed(lambda: eng.distance(tq, tt, res.query_rows, res.target_rows, stranded=True, signed=True))
        emit("C4 DISTANCE over the k=5 pairs", len(res), ms, 29 * len(res))
        del tq, tt, res
        eng.trim()
    T = st.reads(n5, 20260109, dev)
    t = table(T)
    del T
    if world == 1:
        ms, ids = timed(lambda: eng.cluster(t), reps=3, warm=1)
        emit(f"C5 CLUSTER {n5} reads", n5, ms, 16 * n5, clusters_max_id=int(ids.max()))
        del ids
        ms, m = timed(lambda: eng.merge(t), reps=3, warm=1)
        emit(f"C5 MERGE {n5} reads", n5, ms, 12 * n5 + 12 * len(m), regions=len(m))
        assert int(m.count.sum()) == n5
    else:
        from giql_b200 import sharding

        ops = sharding.EngineOps(eng)

        def max_over_ranks(ms):
            v = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return float(v.item())

        def sum_over_ranks(*vals):
            v = torch.tensor(list(vals), dtype=torch.int64, device=dev)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            return [int(x) for x in v.tolist()]

        dist.barrier()
        ms, ids = timed(lambda: sharding.sharded_cluster(ops, rank, world, t), reps=3, warm=1)
        ms = max_over_ranks(ms)
        owned, = sum_over_ranks(int((ids > 0).sum()))
        ms_m, regs = timed(lambda: sharding.sharded_merge(ops, rank, world, t), reps=3, warm=1)
        ms_m = max_over_ranks(ms_m)
        n_regions, n_rows = sum_over_ranks(int(regs["start"].numel()), int(regs["count"].sum()))
        # the single-GPU answer on rank 0's copy of the same table, for the totals
        ref = [0, 0, 0]
        if rank == 0:
            m1 = eng.merge(t)
            ref = [len(m1), int(m1.count.sum()), int(eng.cluster(t).max())]
            del m1
        r = torch.tensor(ref, dtype=torch.int64, device=dev)
        dist.broadcast(r, 0)
        ok = owned == n5 and n_rows == n5 and n_regions == int(r[0])
        emit(f"C5 CLUSTER {n5} reads, genome-range sharded over {world} GPUs (NCCL)", n5, ms, 16 * n5, rows_owned=owned, totals_match_single_gpu=ok)
        emit(f"C5 MERGE {n5} reads, genome-range sharded over {world} GPUs (NCCL)", n5, ms_m, 12 * n5 + 12 * n_regions, regions=n_regions,
             single_gpu_regions=int(r[0]), totals_match_single_gpu=ok)
        assert ok, f"sharded CLUSTER/MERGE totals differ from the single-GPU answer: rows {owned}/{n_rows} of {n5}, regions {n_regions} vs {int(r[0])}"
    del t
    eng.trim()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    import giql_b200
    from giql_b200.engine import DeviceTable

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    # host buffers NUMA-local to this rank's GPU (first touch); the CPU baseline leg gets the full core set back
    prev_aff, new_aff = giql_b200.bind_host_to_gpu(local) if not args.no_bind else (os.sched_getaffinity(0), None)
    log(f"[bench] rank {rank}: cpu affinity {'unchanged (' + str(len(prev_aff)) + ' cores)' if new_aff is None else str(len(prev_aff)) + ' -> ' + str(len(new_aff)) + ' GPU-local cores'}")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = giql_b200.Engine(local)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    A, B = make_workload(args.rows)
    n_l = n_r = args.rows

    def pinned(a):
        return torch.from_numpy(a).pin_memory()

    host = {k: {c: pinned(T[c]) for c in ("offs", "bytes", "start", "end")} for k, T in (("l", A), ("r", B))}
    dev = {k: {c: t.to(eng.device) for c, t in cols.items()} for k, cols in host.items()}
    torch.cuda.synchronize()
    h2d_bytes = sum(t.numel() * t.element_size() for cols in host.values() for t in cols.values()) // max(world, 1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=eng.device)   # > 126 MB L2

    if world > 1:
        from giql_b200 import sharding

        shard = sharding.GenomeRangeShard(eng, rank, world)
    else:
        shard = None

    def resident_tables():
        # fresh DeviceTable objects over the SAME device buffers, so the dictionary is re-encoded each step
        mk = lambda c: DeviceTable(eng, int(c["start"].numel()), None, c["start"], c["end"], None, c["offs"], c["bytes"])
        return mk(dev["l"]), mk(dev["r"])

    def step_device():
        tl, tr = resident_tables()
        if shard is not None:
            return shard.intersects(tl, tr)
        return eng.intersects(tl, tr, mode=args.join_mode)

    def step_e2e(out_l_host, out_r_host):
        if shard is not None:
            # every rank uploads 1/N of the rows over its own PCIe link; NCCL all-gathers the blocks over NVLink
            tl = sharding.upload_replicated(eng, host["l"], rank, world)
            tr = sharding.upload_replicated(eng, host["r"], rank, world)
        else:
            tl = eng.table_from_buffers(host["l"]["offs"], host["l"]["bytes"], host["l"]["start"], host["l"]["end"])
            tr = eng.table_from_buffers(host["r"]["offs"], host["r"]["bytes"], host["r"]["start"], host["r"]["end"])
        res = shard.intersects(tl, tr) if shard is not None else eng.intersects(tl, tr, mode=args.join_mode)
        n = len(res)
        out_l_host[:n].copy_(res.left_rows, non_blocking=True)
        out_r_host[:n].copy_(res.right_rows, non_blocking=True)
        torch.cuda.synchronize()
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the pinned result buffers) ----
    pairs_local = 0
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        pairs_local = len(step_device())
    torch.cuda.synchronize()
    out_l_host = torch.empty(max(pairs_local, 1), dtype=torch.int32).pin_memory()
    out_r_host = torch.empty(max(pairs_local, 1), dtype=torch.int32).pin_memory()

    # ---- timed: device-resident steps, one CUDA-event pair per step, L2 flushed between steps ----
    def timed_pass():
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        with ClockSampler(local) as clk:
            for i in range(args.steps):
                flush.fill_(i & 0xFF)
                starts[i].record(eng.stream)
                res = step_device()
                stops[i].record(eng.stream)
                del res
            torch.cuda.synchronize()
        barrier()
        return [s.elapsed_time(e) for s, e in zip(starts, stops)], clk

    # pass 1 produces `value`; pass 2 repeats the same K steps with the library's per-kernel event spans switched on
    # and produces `phases` / `roofline` (the spans cost ~5 % of a step, so they stay off in pass 1)
    k0 = eng.kernels_total()
    step_ms, clocks = timed_pass()
    launches = eng.kernels_total() - k0
    big_tiles = eng.stats().get("big_tiles")
    eng.profile(True)
    eng.profile_read()
    prof_step_ms, _ = timed_pass()
    prof = eng.profile_read()
    eng.profile(False)
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=eng.device)
    pairs_t = torch.tensor([pairs_local], dtype=torch.int64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs_t, op=dist.ReduceOp.SUM)
    total_ms = float(t_dev.item())
    pairs = int(pairs_t.item())
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step / 1e3)

    # ---- timed: end to end from pinned host buffers through the public API ----
    # (1) one synchronous call per step: upload, join, download, synchronize (the latency of a single request)
    for _ in range(2):
        step_e2e(out_l_host, out_r_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e(out_l_host, out_r_host)
    barrier()
    e2e_sync_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    # (2) the same K steps through Engine.intersects_batches (GenomeRangeShard.intersects_batches when sharded): every
    # step still uploads its inputs and downloads its pairs, but step i+1's upload and step i's download overlap the
    # compute (PCIe is full duplex).
    e2e_mode, e2e_local_s, e2e_pipelined_ms = "synchronous call per step", e2e_sync_ms * args.steps / 1e3, None
    if not args.no_pipeline:
        try:
            batch = (host["l"], host["r"])
            run = (lambda k: eng.intersects_batches(iter([batch] * k), mode=args.join_mode)) if shard is None else \
                  (lambda k: shard.intersects_batches(iter([batch] * k)))
            for n_b, _, _ in run(max(3, min(args.steps, 10))):          # untimed warm-up
                assert n_b == pairs_local
            barrier()
            t0 = time.perf_counter()
            seen = 0
            for n_b, _, _ in run(args.steps):
                seen += n_b
            barrier()
            assert seen == pairs_local * args.steps
            pipelined_s = time.perf_counter() - t0
            if world > 1:                                               # one decision for all ranks
                ps = torch.tensor([pipelined_s, e2e_local_s], dtype=torch.float64, device=eng.device)
                dist.all_reduce(ps, op=dist.ReduceOp.MAX)
                pipelined_s, sync_s = float(ps[0].item()), float(ps[1].item())
            else:
                sync_s = e2e_local_s
            e2e_pipelined_ms = 1e3 * pipelined_s / args.steps
            if pipelined_s > sync_s:                                    # a user would pick the faster schedule; both are reported
                raise RuntimeError(f"pipelined schedule slower than synchronous calls on this box ({e2e_pipelined_ms:.2f} ms vs {1e3 * sync_s / args.steps:.2f} ms)")
            e2e_local_s = pipelined_s
            e2e_mode = "pipelined over steps, depth 2 (upload of step i+1 and download of step i overlap compute; " + \
                       ("Engine.intersects_batches)" if shard is None else "GenomeRangeShard.intersects_batches: 1/N upload per rank + NCCL all-gather)")
        except Exception as exc:   # noqa: BLE001 — keep the synchronous number rather than lose the bench line
            log(f"[bench] pipelined e2e schedule not used ({type(exc).__name__}: {exc}); reporting the synchronous schedule")
            if world > 1 and not isinstance(exc, RuntimeError):
                raise                                                   # ranks would disagree about the collectives still to come
    e2e_s = torch.tensor([e2e_local_s], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(e2e_s.item()) / args.steps
    d2h_bytes = 8 * pairs_local

    # ---- the other BASELINE configs (all ranks take part in the sharded CLUSTER / MERGE leg) ----
    others = None
    if not args.no_other:
        del dev, flush
        eng.trim()
        torch.cuda.empty_cache()
        try:
            others = other_configs(eng, rank, world, hbm_peak, args.other_scale)
        except Exception as exc:   # noqa: BLE001 — the headline line must survive a failure here; the failure is reported in it
            if world > 1:
                raise
            others = [{"error": f"{type(exc).__name__}: {exc}"}]
            log(f"[bench] other_configs failed: {others[0]['error']}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel group (per-launch, live CUDA events on the engine's stream) ----
    # digit passes per table as launched (three when the one-call path took the short-window sort, four otherwise)
    sort_passes = prof["sort_pass"][1] / args.steps / 2 if prof.get("sort_pass", (0, 0))[1] > 0 else 4
    phases = {}
    for name, (ms, n_launch) in prof.items():
        if n_launch <= 0:
            continue
        per_step_ms = ms / args.steps
        ab = algorithmic_bytes(name, n_l + n_r, (n_l + n_r) // world, pairs_local, sort_passes)
        phases[name] = {"ms_per_step": round(per_step_ms, 4), "launches_per_step": n_launch / args.steps,
                        "algorithmic_bytes_per_step": ab,
                        "gbs": round(ab / (per_step_ms * 1e6), 1) if per_step_ms > 0 else None}
    dom = max(phases, key=lambda k: phases[k]["ms_per_step"])
    d = phases[dom]
    per_launch_ms = d["ms_per_step"] / d["launches_per_step"]
    per_launch_bytes = d["algorithmic_bytes_per_step"] / d["launches_per_step"]
    achieved = per_launch_bytes / (per_launch_ms * 1e6)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": None, "peak_source": peak_src,
                "ms_per_launch": round(per_launch_ms, 4), "algorithmic_bytes_per_launch": int(per_launch_bytes)}
    # DRAM bytes per launch cannot be measured without a profiler: the figure comes from the committed ncu capture of the
    # same command at N = 1 (profiles/traffic.json names the capture), and is left null for sharded runs, whose launches
    # move a different number of rows
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists() and world == 1:
        try:
            tj = json.loads(traffic_file.read_text())
            roofline["traffic"] = tj.get(dom)
            roofline["traffic_source"] = tj.get("_source", {}).get(dom)
        except Exception:
            pass
    join_ms = sum(phases[p]["ms_per_step"] for p in ("join_count", "join_scan", "join_emit") if p in phases)
    join_bytes = 12 * ((n_l + n_r) // world) + 8 * pairs_local
    join_roofline = {"ms": round(join_ms, 4), "algorithmic_bytes": join_bytes,
                     "achieved": round(join_bytes / (join_ms * 1e6), 1) if join_ms else None, "unit": "GB/s",
                     "frac": round(join_bytes / (join_ms * 1e6) / hbm_peak, 4) if join_ms else None,
                     "frac_whole_step": round(join_bytes / (ms_per_step * 1e6) / hbm_peak, 4)}

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    os.sched_setaffinity(0, prev_aff)
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle

        c_oracle.build()
        threads = c_oracle.default_threads()
        cpu_join(A, B, threads)   # sizes the output and warms the page cache
        total, dt = cpu_join(A, B, threads)
        assert total == pairs, f"CPU restatement found {total} pairs, GPU {pairs}"
        cpu = {"value": total / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"full workload, 1 run after 1 warm-up ({dt:.2f}s): C/OpenMP restatement, NOT DuckDB (duckdb/sqlglot absent)"}
        duck = duckdb_baseline(A, B, threads)
        if duck is not None:
            assert duck["pairs"] == pairs, f"DuckDB found {duck['pairs']} pairs, GPU {pairs}"
            cpu = {k: duck[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["port_value"] = total / dt   # the C restatement beside it

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "ms_per_step_min_median_max": [round(min(step_ms), 4), round(float(np.median(step_ms)), 4), round(max(step_ms), 4)],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.rows), "pairs": pairs, "rows_left": n_l, "rows_right": n_r, "join_tiles_delegated_to_big_kernel": big_tiles,
                   "sharding": "single GPU" if world == 1 else f"{world} contiguous genome ranges balanced on left interval count; "
                               "straddling right intervals replicated; owner = rank of the left start",
                   "l2": "inputs (340 MB) exceed the 126 MB L2 and a 512 MB buffer is rewritten between timed steps",
                   "timing": "value: K steps, CUDA events per step, no instrumentation; phases/roofline: the same K steps repeated with "
                             "CUDA-event spans around every kernel group (ms_per_step_with_spans = %.4f)" % (sum(prof_step_ms) / len(prof_step_ms)),
                   "step": "dict_encode + layout + radix-sort index x2 (upper-bits digit passes + one window-finishing pass that also gathers the ends) + join (" + args.join_mode + ": the warm-up call sizes the pair buffer "
                           "with count+scan+emit; timed calls run the fused single-pass count/offset/emit kernel into a buffer of 1.25x "
                           "the previous total, overflow checked on the device)"},
        "e2e": {"value": pairs / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "ms_per_step_synchronous": e2e_sync_ms, "ms_per_step_pipelined": e2e_pipelined_ms, "schedule": e2e_mode,
                "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes,
                "note": "per-rank bytes" + ("" if world == 1 else "; each rank uploads 1/N of the rows, NCCL all-gathers the blocks over NVLink; "
                                            "each rank reads back its own pairs")},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "join_roofline": join_roofline,
        "phases": phases,
        "clocks": clocks.summary(),
        "cpu_baseline": cpu,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_DEFAULT, help="rows per table (default: the 10M of BASELINE configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE configs (C3 / C4 / C5)")
    ap.add_argument("--other-scale", type=float, default=1.0, help="row-count scale of the other BASELINE configs")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the GPU-local CPU cores")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: one synchronous call per step instead of Engine.intersects_batches")
    ap.add_argument("--join-mode", default="auto", choices=["auto", "two_pass", "single_pass"],
                    help="auto: count+scan+emit on the first call, then the fused single-pass join into a buffer sized from the "
                         "engine's running estimate (overflow is detected on the device and retried exactly)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner
    # there) are pointed at stderr for the duration of the run, and the line goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
