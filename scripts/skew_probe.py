"""One genome-spanning row among 1M: time of every operator with and without it (the skew hazard of SURVEY §7).
usage: python scripts/skew_probe.py [rows]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import giql_b200
from giql_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
eng = giql_b200.Engine(0)


def tables(spanning: bool):
    A, B = synth.peaks(n, 51), synth.peaks(n, 52)
    if spanning:                      # chr1 (code 0) from end to end, once in each table
        for T in (A, B):
            T["start"][0], T["end"][0], T["code"][0] = 0, int(synth.LENGTHS[0]) - 1, 0
    ta = eng.table_from_codes(A["code"], A["start"], A["end"], A["names"], strand=A["strand_code"])
    tb = eng.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary, strand=B["strand_code"])
    return ta, tb


def timed(fn, limit_s=20.0):
    for _ in range(3):        # the first calls of a new shape grow the scratch cache (cudaMalloc), which is not what is measured
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 and time.perf_counter() - t0 < limit_s:
        fn()
        torch.cuda.synchronize()
        reps += 1
    return 1e3 * (time.perf_counter() - t0) / max(reps, 1)


reps_note = "3 reps after three warm-up calls"
ops = {
    "intersects": lambda a, b: eng.intersects(a, b),
    "semi": lambda a, b: eng.semi(a, b),
    "count_overlaps": lambda a, b: eng.count_overlaps(a, b),
    "nearest k=1": lambda a, b: eng.nearest(a, b, k=1),
    "nearest k=5": lambda a, b: eng.nearest(a, b, k=5),
    "nearest k=5 stranded": lambda a, b: eng.nearest(a, b, k=5, stranded=True, signed=True),
    "cluster": lambda a, b: eng.cluster(a),
    "merge": lambda a, b: eng.merge(a),
    "disjoin self": lambda a, b: eng.disjoin(a),
    "disjoin ref": lambda a, b: eng.disjoin(a, b),
    "contains": lambda a, b: eng.contains(a, b),
    "within": lambda a, b: eng.within(a, b),
}
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
base, skew = tables(False), tables(True)
for name, fn in ops.items():
    if only and name not in only:
        continue
    t_u = timed(lambda: fn(*base))
    t_s = timed(lambda: fn(*skew))
    print(f"{name:24s} uniform {t_u:9.3f} ms   with one chromosome-spanning row {t_s:9.3f} ms   x{t_s / t_u:6.2f}", flush=True)
