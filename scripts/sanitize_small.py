"""Small end-to-end exercise of every operator for compute-sanitizer (racecheck / memcheck), a few 10^4 rows each:
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
The fused join is the interesting one: its output staging window aliases the start / end windows across barriers and two
kernels spin on look-back words."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import giql_b200
from giql_b200 import synth

eng = giql_b200.Engine(0)
A, B = synth.peaks(60_000, 1), synth.peaks(50_000, 2)
G = synth.genes(20_000, 3)
ta = eng.table_from_codes(A["code"], A["start"], A["end"], A["names"], strand=A["strand_code"])
tb = eng.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary, strand=B["strand_code"])
tg = eng.table_from_codes(G["code"], G["start"], G["end"], ta.dictionary, strand=G["strand_code"])
n2 = len(eng.intersects(ta, tb, mode="two_pass"))
for _ in range(2):
    n1 = len(eng.intersects(ta, tb, mode="single_pass"))      # the second call runs the fused kernel into a sized buffer
assert n1 == n2
S = synth.with_strings(A)
sa = eng.table({"chrom": S["chrom"], "start": S["start"], "end": S["end"]})
sb = eng.table({"chrom": synth.with_strings(B)["chrom"], "start": B["start"], "end": B["end"]})
assert len(eng.intersects(sa, sb)) == n1                       # one-call pipeline with the device dictionary
for g in range(2):
    eng.intersects_shard(sa, sb, 2, g)
assert int(eng.count_overlaps(ta, tb).sum()) == n1
eng.semi(ta, tb), eng.anti(ta, tb)
eng.nearest(ta, tg, k=3, stranded=True, signed=True)
eng.nearest(ta, tg, k=1)
eng.cluster(ta), eng.cluster(ta, distance=100, stranded=True)
eng.merge(ta), eng.merge(ta, stranded=True)
eng.disjoin(tg, ta), eng.disjoin(ta)
eng.contains(tg, ta), eng.within(ta, tg)
eng.sync()
print(f"sanitize_small ok: {n1} pairs")
