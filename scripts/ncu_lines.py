"""Per CUDA source line: executed warp instructions and stall samples, from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K --launch-count 1`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
fname, hdr, agg = None, None, collections.OrderedDict()
src = {}
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {h: i for i, h in enumerate(hdr)}
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        cur = (fname, int(r[0]))
        src[cur] = r[1].strip()
        agg.setdefault(cur, [0, 0])
    elif cur and r[2] not in ("-", "..."):
        try:
            agg[cur][0] += int(r[ii] or 0)
            agg[cur][1] += int(r[si] or 0)
        except ValueError:
            pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instr {ti}, samples {ts}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{v[0] / ti:6.3f} ins {v[1] / ts:6.3f} smp  {k[0]}:{k[1]:<5d} {src[k][:100]}")
