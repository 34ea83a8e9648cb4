#!/bin/bash
# A/B bench lines of library variants in one gpurun call.
# usage: scripts/gpu_ab.sh <tag> "<name>=<ENV1=..,ENV2=..>" ...   (name "default" = the in-tree build; otherwise gpurun_variants/<name>/)
# a config "name+test" also runs the GPU parity suite under it.
set -u
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}; [ "$envs" = "$spec" ] && envs=""
  runtest=0; case "$name" in *+test) runtest=1; name=${name%+test};; esac
  lib=${name%%@*}
  E=()
  [ "$lib" != "default" ] && E+=("GIQLB_LIB=$PWD/gpurun_variants/$lib/libgiqlb200.so")
  IFS=',' read -ra KV <<< "$envs"; for kv in "${KV[@]}"; do [ -n "$kv" ] && E+=("$kv"); done
  label=$(echo "$name" | tr '@' '_')
  if [ $runtest = 1 ]; then
    env "${E[@]}" timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}_$label.log 2>&1; echo "[$label] pytest rc=$? $(tail -1 gpurun_out/pytest_${TAG}_$label.log)"
  fi
  env "${E[@]}" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other > gpurun_out/ab_${TAG}_$label.json 2> gpurun_out/ab_${TAG}_$label.err; rc=$?
  python - "$label" gpurun_out/ab_${TAG}_$label.json $rc <<'PY'
import json, sys
label, path, rc = sys.argv[1:4]
try:
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    ph = d.get("phases", {})
    print(f"[{label}] rc={rc} step={d['ms_per_step']:.4f} ms  join={d.get('join_roofline', {}).get('ms')} frac={d.get('join_roofline', {}).get('frac')}  " +
          " ".join(f"{k}={v['ms_per_step']:.4f}" for k, v in ph.items()) + f"  e2e={d['e2e']['ms_per_step']:.2f}")
except Exception as e:
    print(f"[{label}] rc={rc} no line: {e}")
PY
done
