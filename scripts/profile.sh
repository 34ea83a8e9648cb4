#!/bin/bash
# ncu evidence for bench.py (run under gpurun, one GPU).  usage: scripts/profile.sh <tag> [kernel-regex] [skip] [count]
set -u
TAG=${1:-r01}
REGEX=${2:-'join_emit|join_count|rs_onesweep|dict_insert|scan_lookback'}
SKIP=${3:-40}
COUNT=${4:-14}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-other"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s $SKIP -c $COUNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
