#!/bin/bash
# builds and runs the CUB comparison bar on the GPU box; output -> gpurun_out/cub_bar.json
set -e
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/cub_sort_bar cub_sort_bar.cu
/tmp/cub_sort_bar 10000000 | tee ../../gpurun_out/cub_bar.json
