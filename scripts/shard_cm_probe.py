"""Per-phase device time of ONE rank's share of sharded CLUSTER / MERGE (1 GPU playing rank 0 of `world`; the all-gathers are
replaced by the values a single process can compute).  usage: python scripts/shard_cm_probe.py [rows] [world]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import giql_b200
from giql_b200 import sharding, synth_torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = giql_b200.Engine(0)
T = synth_torch.reads(n, 20260109, eng.device)
t = eng.table_from_codes(T["code"], T["start"], T["end"], T["names"])
del T
ops = sharding.EngineOps(eng)


def step(label, fn, reps=2):
    fn()
    eng.profile(True)
    eng.profile_read()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ph = eng.profile_read()
    eng.profile(False)
    print(f"{label}: {wall:.2f} ms wall; " + ", ".join(f"{k} {v[0] / reps:.3f}ms x{v[1] // reps}" for k, v in ph.items() if v[1]), flush=True)
    return out


def cluster_rank0():
    plan = ops.cm_plan(t, world)
    try:
        ops.cm_max_end(plan, 0)
        counts = ops.cm_cluster_counts(plan, 0, 0, 0)
        return ops.cm_cluster_ids(plan, 0, 0, 0, [0] * len(counts))
    finally:
        ops.cm_release(plan)


def merge_rank0():
    plan = ops.cm_plan(t, world)
    try:
        ops.cm_max_end(plan, 0)
        return ops.cm_merge(plan, 0, 0, 0)
    finally:
        ops.cm_release(plan)


step(f"sharded CLUSTER rank 0 of {world}, {n} rows", cluster_rank0)
step(f"sharded MERGE rank 0 of {world}, {n} rows", merge_rank0)
