"""Summarise an `ncu --page source --csv` dump (first kernel in the file): stall-reason totals, the hottest SASS
instructions by samples, and executed-instruction totals."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        if data:
            break
        continue
    data.append(r)
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
alls = sum(tot.values()) or 1
print("stall totals:", ", ".join(f"{k[6:]}={v / alls:.2f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
si, ii = col["# Samples"], col["Instructions Executed"]
tots = sum(int(r[si] or 0) for r in data) or 1
toti = sum(int(r[ii] or 0) for r in data) or 1
print(f"SASS instructions: {len(data)}  executed warp-instr: {toti}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:n]:
    best = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"{int(r[si]) / tots:6.3f} smp {int(r[ii]) / toti:6.3f} ins  {r[col['Source']].strip()[:80]:80s} {best[6:]}")
