"""GPU probe: rank 0 of a `world`-way sharded join on one GPU (for an ncu launch list of the rank-local kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import giql_b200
from giql_b200.engine import DeviceTable
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
A, B = bench.make_workload(10_000_000)
eng = giql_b200.Engine(0)
dev = {k: {c: torch.from_numpy(T[c]).to(eng.device) for c in ("offs", "bytes", "start", "end")} for k, T in (("l", A), ("r", B))}
mk = lambda c: DeviceTable(eng, int(c["start"].numel()), None, c["start"], c["end"], None, c["offs"], c["bytes"])
for _ in range(4):
    r = eng.intersects_shard(mk(dev["l"]), mk(dev["r"]), world, 0)
torch.cuda.synchronize()
print("pairs on rank 0:", len(r))
