"""Per-phase cycle breakdown of the fused join (a -DGIQLB_JOIN_DEBUG_STATS build selected with GIQLB_LIB)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import giql_b200
from giql_b200 import synth, _cabi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
eng = giql_b200.Engine(0)
A, B = synth.peaks(rows, 20260103), synth.peaks(rows, 20260104)
ta = eng.table_from_codes(A["code"], A["start"], A["end"], A["names"])
tb = eng.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary)
lib = ctypes.CDLL(str(_cabi.LIB_PATH))
out = (ctypes.c_ulonglong * 32)()
for it in range(3):
    eng.intersects(ta, tb, mode="single_pass")
    lib.giqlb_debug_join_stats(out, 1)
v = np.array(list(out), dtype=np.float64)
n = v[0]
names = ["tiles", "ticket+geom", "copy wait", "runs", "scan", "table", "expand(t0)", "barrier after expand(t0)", "flush", "lookback(warp)", "lb steps", "lb polls", "whole tile"]
for i, nm in enumerate(names):
    print(f"{nm:28s} {v[i] / n:10.1f} per tile" + (f"  ({100 * v[i] / v[12]:.1f} % of tile)" if 1 <= i <= 9 else ""))
