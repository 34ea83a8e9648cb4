"""Throughput of the other BASELINE.json configs on one B200 (rows/s, inputs resident in HBM, CUDA events):

  C1  INTERSECTS inner join 100k x 100k                      (configs[0], the reference's CPU-runnable case)
  C3  SEMI and ANTI, 50M query x 1M target                   (configs[2])
  C4  NEAREST k=1 / k=5 stranded + DISTANCE, 10M x 2M        (configs[3])
  C5  MERGE and CLUSTER over 200M reads-as-intervals         (configs[4], single GPU)

These are parity-test shapes, not the bench.py line; the numbers go into DESIGN.md.  usage:
    python scripts/bench_configs.py [--scale 1.0] [--only c3,c4] > gpurun_out/configs.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import giql_b200
from giql_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--only", default="c1,c3,c4,c5,bed,strand")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
only = set(args.only.split(","))
eng = giql_b200.Engine(0)
PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0) \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0


def table(T, strand=False, dictionary=None):
    return eng.table_from_codes(T["code"], T["start"], T["end"], dictionary or T["names"], strand=T["strand_code"] if strand else None)


def timed(fn, reps=args.reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


def emit(name, rows, ms, alg_bytes, **extra):
    line = {"config": name, "rows": rows, "ms": round(ms, 3), "rows_per_s": rows / (ms / 1e3), "algorithmic_bytes": alg_bytes,
            "gbs": round(alg_bytes / (ms * 1e6), 1), "frac_of_measured_hbm": round(alg_bytes / (ms * 1e6) / PEAK, 4)}
    line.update(extra)
    print(json.dumps(line), flush=True)


if "c1" in only:
    A, B = synth.peaks(100_000, 20260101), synth.peaks(100_000, 20260102)
    ta = table(A)
    tb = table(B, dictionary=ta.dictionary)
    ms, res = timed(lambda: eng.intersects(ta, tb))
    emit("C1 INTERSECTS 100k x 100k", 200_000, ms, 12 * 200_000 + 8 * len(res), pairs=len(res))

if "c3" in only:
    nq, nt = int(50_000_000 * args.scale), int(1_000_000 * args.scale)
    t0 = time.time()
    Q, T = synth.peaks(nq, 20260105), synth.peaks(nt, 20260106)
    tq = table(Q)
    tt = table(T, dictionary=tq.dictionary)
    print(f"# C3 generated in {time.time() - t0:.0f}s", file=sys.stderr)
    ms, rows = timed(lambda: eng.semi(tq, tt))
    emit("C3 SEMI 50M x 1M", nq, ms, 12 * (nq + nt) + 4 * int(rows.numel()), kept=int(rows.numel()))
    ms, rows_a = timed(lambda: eng.anti(tq, tt))
    emit("C3 ANTI 50M x 1M", nq, ms, 12 * (nq + nt) + 4 * int(rows_a.numel()), kept=int(rows_a.numel()))
    assert int(rows.numel()) + int(rows_a.numel()) == nq
    del tq, tt, Q, T

if "c4" in only:
    nq, nt = int(10_000_000 * args.scale), int(2_000_000 * args.scale)
    Q, T = synth.peaks(nq, 20260107), synth.genes(nt, 20260108)
    tq = table(Q, strand=True)
    tt = table(T, strand=True, dictionary=tq.dictionary)
    for k in (1, 5):
        ms, res = timed(lambda: eng.nearest(tq, tt, k=k, stranded=True, signed=True))
        emit(f"C4 NEAREST k={k} stranded 10M x 2M", nq, ms, 13 * (nq + nt) + 12 * len(res), emitted=len(res))
    ms, _ = timed(lambda: eng.distance(tq, tt, res.query_rows.long(), res.target_rows.long(), stranded=True, signed=True))
    emit("C4 DISTANCE over the k=5 pairs", len(res), ms, 29 * len(res))
    del tq, tt, Q, T, res

if "c5" in only:
    n = int(200_000_000 * args.scale)
    t0 = time.time()
    parts = [synth.reads(n // 4, 20260109 + i) for i in range(4)]      # four shuffled quarters, concatenated on the device
    names = parts[0]["names"]
    code = torch.cat([torch.from_numpy(p["code"]).to(eng.device) for p in parts])
    start = torch.cat([torch.from_numpy(p["start"]).to(eng.device) for p in parts])
    end = torch.cat([torch.from_numpy(p["end"]).to(eng.device) for p in parts])
    del parts
    print(f"# C5 generated in {time.time() - t0:.0f}s", file=sys.stderr)
    t = eng.table_from_codes(code, start, end, names)
    n = t.n
    ms, ids = timed(lambda: eng.cluster(t), reps=3)
    emit("C5 CLUSTER 200M", n, ms, 16 * n, clusters_on_chr1=int(ids.max()))
    ms, m = timed(lambda: eng.merge(t), reps=3)
    emit("C5 MERGE 200M", n, ms, 12 * n + 12 * len(m), regions=len(m))
    assert int(m.count.sum()) == n

if "bed" in only:
    import io
    import pyarrow as pa
    import pyarrow.csv as pc

    n = int(10_000_000 * args.scale)
    T = synth.with_strings(synth.peaks(n, 20260103))
    tab = pa.table({"chrom": pa.array(T["chrom"], type=pa.string()), "start": T["start"], "end": T["end"],
                    "name": pa.array(np.char.add("p", np.arange(n).astype(str))), "score": np.zeros(n, np.int32), "strand": pa.array(T["strand"], type=pa.string())})
    sink = io.BytesIO()
    pc.write_csv(tab, sink, pc.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
    raw = np.frombuffer(sink.getvalue(), dtype=np.uint8).copy()
    text = torch.from_numpy(raw).to(eng.device)
    import ctypes
    from giql_b200.engine import _ptr

    def parse_resident():
        nl = ctypes.c_int64()
        eng._check(eng.lib.giqlb_bed_count(eng.ctx, _ptr(text), text.numel(), ctypes.byref(nl)))
        m = nl.value
        cols = [eng._empty(m + 1, torch.int32), eng._empty(8 * m + 16, torch.uint8), eng._empty(m, torch.int32), eng._empty(m, torch.int32),
                eng._empty(m + 1, torch.int32), eng._empty(16 * m + 16, torch.uint8), eng._empty(m, torch.int32), eng._empty(m, torch.uint8)]
        a, b = ctypes.c_int64(), ctypes.c_int64()
        eng._check(eng.lib.giqlb_bed_parse(eng.ctx, _ptr(text), text.numel(), m, 6, _ptr(cols[0]), _ptr(cols[1]), 8 * m, _ptr(cols[2]), _ptr(cols[3]),
                                           _ptr(cols[4]), _ptr(cols[5]), 16 * m, _ptr(cols[6]), _ptr(cols[7]), ctypes.byref(a), ctypes.byref(b)))
        return m, a.value, b.value

    ms, (m, cb, nb) = timed(parse_resident)
    assert m == n
    emit("BED6 text parse, 10M lines resident in HBM", n, ms, 3 * len(raw) + 21 * n + cb + nb, text_bytes=len(raw), text_gbs=round(len(raw) / (ms * 1e6), 1))
    t0 = time.perf_counter()
    t = eng.read_bed(raw)
    torch.cuda.synchronize()
    emit("BED6 read_bed from host bytes (pageable H2D + parse)", n, (time.perf_counter() - t0) * 1e3, 3 * len(raw) + 21 * n + cb + nb, text_bytes=len(raw))

if "strand" in only:
    nq = int(10_000_000 * args.scale)
    A, B = synth.peaks(nq, 20260103), synth.peaks(nq, 20260104)
    ta = table(A, strand=True)
    tb = table(B, strand=True, dictionary=ta.dictionary)
    ms, res = timed(lambda: eng.intersects(ta, tb))
    ms2, kept = timed(lambda: eng.filter_pairs(res, ta.strand, tb.strand, "=", strand=True))
    emit("same-strand residual over the C2 pairs (giqlb_pairs_filter)", len(res), ms2, 10 * len(res) + 8 * len(kept), kept=len(kept), join_ms=round(ms, 3))
