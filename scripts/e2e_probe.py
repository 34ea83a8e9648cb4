"""Where the sharded end-to-end step spends its wall-clock time (run under torchrun): upload + all-gather, join, download.
usage: python -m torch.distributed.run --nproc-per-node N scripts/e2e_probe.py"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import bench
import giql_b200
from giql_b200 import sharding

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = giql_b200.Engine(local)
A, B = bench.make_workload(10_000_000)
pin = lambda a: torch.from_numpy(a).pin_memory()
host = {k: {c: pin(T[c]) for c in ("offs", "bytes", "start", "end")} for k, T in (("l", A), ("r", B))}
shard = sharding.GenomeRangeShard(eng, rank, world) if world > 1 else None
out_l = out_r = None
acc = {"upload": 0.0, "join": 0.0, "download": 0.0}
for it in range(13):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tl = sharding.upload_replicated(eng, host["l"], rank, world)
    tr = sharding.upload_replicated(eng, host["r"], rank, world)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    res = shard.intersects(tl, tr) if shard else eng.intersects(tl, tr)
    n = len(res)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if out_l is None:
        out_l, out_r = torch.empty(int(n * 1.3), dtype=torch.int32).pin_memory(), torch.empty(int(n * 1.3), dtype=torch.int32).pin_memory()
    out_l[:n].copy_(res.left_rows, non_blocking=True)
    out_r[:n].copy_(res.right_rows, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 3:
        acc["upload"] += t1 - t0
        acc["join"] += t2 - t1
        acc["download"] += t3 - t2
print(f"rank {rank}/{world}: " + ", ".join(f"{k} {1e3 * v / 10:.2f} ms" for k, v in acc.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
