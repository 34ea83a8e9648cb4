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
ap.add_argument("--only", default="c1,c3,c4,c5")
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
