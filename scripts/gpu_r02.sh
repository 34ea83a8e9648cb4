#!/bin/bash
# One gpurun call: parity suite, bench line, launch list and one full capture of the named kernels.
# usage: scripts/gpu_r02.sh <tag> [kernel-regex] [skip] [count]
set -u
TAG=${1:-r02a}
REGEX=${2:-'join_fused'}
SKIP=${3:-3}
COUNT=${4:-1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-other"
timeout 300 $CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s $SKIP -c $COUNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
python scripts/show_bench.py gpurun_out/bench_$TAG.json 2>/dev/null | head -40
