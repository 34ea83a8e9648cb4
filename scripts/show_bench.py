"""Pretty-print a bench.py JSON line: headline numbers and the per-phase table."""
import json, sys
for line in open(sys.argv[1]):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(f"n_gpus={d['n_gpus']} ms/step={d['ms_per_step']:.3f} value={d['value'] / 1e9:.2f} G{d['unit']} e2e={d['e2e']['ms_per_step']:.2f} ms "
          f"({d['e2e']['value'] / 1e9:.2f} G) launches={d['gpu_launches']} clocks={d.get('clocks')}")
    print("step min/median/max:", d.get("ms_per_step_min_median_max"))
    print("roofline:", d.get("roofline"))
    print("join_roofline:", d.get("join_roofline"))
    tot = 0
    for k, v in d.get("phases", {}).items():
        tot += v["ms_per_step"]
        print(f"  {k:12s} {v['ms_per_step']:8.4f} ms  x{v['launches_per_step']:<4.0f} {v['gbs']} GB/s")
    print(f"  sum of phases {tot:.3f} ms")
    if d.get("cpu_baseline"):
        print("cpu:", d["cpu_baseline"])
