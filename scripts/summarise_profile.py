"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, plain_TAG.json} into the tracked summaries under profiles/.

usage: python scripts/summarise_profile.py TAG     (reads gpurun_out/, writes profiles/TAG_*.{txt,csv,json})
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
src, dst = ROOT / "gpurun_out", ROOT / "profiles"
dst.mkdir(exist_ok=True)

# 1. launch list: per-kernel count, total, share
rows = list(csv.reader(open(src / f"launches_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values())
with open(dst / f"{tag}_launches.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: bench.py --steps 3 --warmup 3 --no-cpu ({len(data)} launches)\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's phases, not absolutes\n")
    f.write(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]:10.1f} {v[1] / tot:6.3f} {v[1] / v[0]:8.1f}\n")

# 2. full capture: the metrics the roofline needs, one row per profiled launch
rep = src / f"prof_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, body = rr[0], rr[1], rr[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct"]
    idx = [h.index(w) for w in want if w in h]
    with open(dst / f"{tag}_kernels.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in body:
            row = [r[i] for i in idx]
            row[0] = row[0].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            w.writerow(row)
    # per-kernel DRAM traffic (bytes per launch) for bench.py's roofline.traffic
    traffic = collections.defaultdict(list)
    bi, wi, ni = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Kernel Name")
    def to_bytes(v, u):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    for r in body:
        name = r[ni].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0]
        traffic[name].append(to_bytes(r[bi], units[bi]) + to_bytes(r[wi], units[wi]))
    (dst / f"{tag}_traffic.json").write_text(json.dumps({k: int(sum(v) / len(v)) for k, v in traffic.items()}, indent=1))

plain = src / f"plain_{tag}.json"
if plain.exists():
    (dst / f"{tag}_bench.json").write_text(plain.read_text())
print("wrote", sorted(p.name for p in dst.glob(f"{tag}_*")))
