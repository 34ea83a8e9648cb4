"""CLUSTER and MERGE over N synthetic reads, twice each (the second call runs with warm scratch) — for `ncu --metrics
gpu__time_duration.sum` launch lists.  usage: python scripts/c5_once.py [rows]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import giql_b200
from giql_b200 import synth_torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
eng = giql_b200.Engine(0)
T = synth_torch.reads(n, 20260109, eng.device)
t = eng.table_from_codes(T["code"], T["start"], T["end"], T["names"])
for _ in range(2):
    ids = eng.cluster(t)
    m = eng.merge(t)
torch.cuda.synchronize()
print("clusters", int(ids.max()), "regions", len(m))
