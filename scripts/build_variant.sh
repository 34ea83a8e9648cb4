#!/bin/bash
# Builds an A/B variant of libgiqlb200.so into gpurun_variants/<name>/ (git-ignored, travels with gpurun).
# usage: scripts/build_variant.sh <name> "<extra nvcc -D flags>"     then run with GIQLB_LIB=gpurun_variants/<name>/libgiqlb200.so
set -eu
NAME=$1; FLAGS=${2:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_variants/$NAME
mkdir -p $OUT/obj
cd $ROOT/giql_b200/csrc
for f in *.cu; do
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $FLAGS -c $f -o $OUT/obj/${f%.cu}.o 2> $OUT/obj/${f%.cu}.log || { cat $OUT/obj/${f%.cu}.log; exit 1; } ) &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libgiqlb200.so $OUT/obj/*.o -lcudart
echo "built $OUT/libgiqlb200.so ($FLAGS)"
