"""GPU probe: step time of the C2 join with and without the per-phase event spans (how much do the spans cost?)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import giql_b200
from giql_b200.engine import DeviceTable
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
A, B = bench.make_workload(rows)
eng = giql_b200.Engine(0)
dev = {k: {c: torch.from_numpy(T[c]).to(eng.device) for c in ("offs", "bytes", "start", "end")} for k, T in (("l", A), ("r", B))}
mk = lambda c: DeviceTable(eng, int(c["start"].numel()), None, c["start"], c["end"], None, c["offs"], c["bytes"])
def step():
    return eng.intersects(mk(dev["l"]), mk(dev["r"]))
for _ in range(4):
    step()
for prof in (False, True, False):
    eng.profile(prof)
    eng.profile_read()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10 * 1e3
    print(f"profile={prof}: device step median {np.median(ts):.3f} ms  min {min(ts):.3f}  wall/step {wall:.3f} ms")
