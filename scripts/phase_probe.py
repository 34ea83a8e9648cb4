"""Per-phase device time (the library's CUDA-event spans) of the non-headline operators at a chosen scale.
usage: python scripts/phase_probe.py [--scale 0.25] [--only c3,c4,c5]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import giql_b200
from giql_b200 import synth_torch

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--only", default="c3,c4,c5")
args = ap.parse_args()
only = set(args.only.split(","))
eng = giql_b200.Engine(0)


class synth:   # the bench's device-side generators (the host ones take minutes at 200 M rows)
    peaks = staticmethod(lambda n, seed: synth_torch.peaks(n, seed, eng.device))
    genes = staticmethod(lambda n, seed: synth_torch.genes(n, seed, eng.device))
    reads = staticmethod(lambda n, seed: synth_torch.reads(n, seed, eng.device))


def table(T, strand=False, dictionary=None):
    return eng.table_from_codes(T["code"], T["start"], T["end"], dictionary or T["names"], strand=T["strand_code"] if strand else None)


def probe(name, fn, reps=3):
    for _ in range(2):
        fn()
    eng.profile(True)
    eng.profile_read()
    for _ in range(reps):
        fn()
    ph = eng.profile_read()
    eng.profile(False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / reps:.3f} ms/call; " + ", ".join(f"{k} {v[0] / reps:.3f}ms x{v[1] // reps}" for k, v in ph.items() if v[1]), flush=True)


if "c3" in only:
    nq, nt = int(50_000_000 * args.scale), int(1_000_000 * args.scale)
    tq = table(synth.peaks(nq, 5))
    tt = table(synth.peaks(nt, 6), dictionary=tq.dictionary)
    probe(f"SEMI {nq} x {nt}", lambda: eng.semi(tq, tt))
    probe(f"count_overlaps {nq} x {nt}", lambda: eng.count_overlaps(tq, tt))
    del tq, tt
if "c4" in only:
    nq, nt = int(10_000_000 * args.scale), int(2_000_000 * args.scale)
    tq = table(synth.peaks(nq, 7), strand=True)
    tt = table(synth.genes(nt, 8), strand=True, dictionary=tq.dictionary)
    for k in (1, 5):
        probe(f"NEAREST k={k} stranded {nq} x {nt}", lambda: eng.nearest(tq, tt, k=k, stranded=True, signed=True))
    probe(f"NEAREST k=1 unstranded {nq} x {nt}", lambda: eng.nearest(tq, tt, k=1))
    del tq, tt
if "c5" in only:
    n = int(200_000_000 * args.scale)
    t = table(synth.reads(n, 9))
    probe(f"CLUSTER {n}", lambda: eng.cluster(t))
    probe(f"MERGE {n}", lambda: eng.merge(t))
    td = table(synth.peaks(n // 10, 10))                    # built outside the timed lambda (round 1 timed the host generator)
    probe(f"DISJOIN {n // 10}", lambda: eng.disjoin(td))
