"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
alls = sum(tot.values())
print("stall totals:", ", ".join(f"{k[6:]}={v / alls:.2f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
si = col["# Samples"]
top = sorted(data, key=lambda r: -int(r[si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
tots = sum(int(r[si] or 0) for r in data)
for r in top:
    best = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"{int(r[si]) / tots:6.3f} {r[col['Source']].strip()[:90]:90s} {best[6:]}")
