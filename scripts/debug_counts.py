"""GPU debug: count_overlaps vs the C oracle at several sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import giql_b200
from giql_b200 import synth
from oracle import c_oracle

eng = giql_b200.Engine(0)
for na, nb in ((1000, 900), (5000, 4096), (5000, 4097), (50_000, 40_000), (100_000, 100_000)):
    A, B = synth.peaks(na, 1), synth.peaks(nb, 2)
    ta = eng.table_from_codes(A["code"], A["start"], A["end"], A["names"])
    tb = eng.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary)
    cnt = eng.count_overlaps(ta, tb).cpu().numpy()
    exp = c_oracle.intersects_counts(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], 24)
    bad = np.nonzero(cnt != exp)[0]
    print(na, nb, "sum", cnt.sum(), exp.sum(), "bad rows", len(bad))
    for i in bad[:5]:
        print("   row", i, "code", A["code"][i], A["start"][i], A["end"][i], "got", cnt[i], "exp", exp[i])
