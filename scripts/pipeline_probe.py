"""Repeatability of the pipelined end-to-end schedule on one box: K batches, repeated, for a few depths."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
import giql_b200

eng = giql_b200.Engine(0)
A, B = bench.make_workload(10_000_000)
pin = lambda a: torch.from_numpy(a).pin_memory()
host = {k: {c: pin(T[c]) for c in ("offs", "bytes", "start", "end")} for k, T in (("l", A), ("r", B))}
batch = (host["l"], host["r"])
for depth in (2, 3):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in eng.intersects_batches(iter([batch] * 10), depth=depth):
            pass
        torch.cuda.synchronize()
        print(f"depth {depth} rep {rep}: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms/batch", flush=True)
