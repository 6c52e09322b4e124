#!/bin/bash
# Build the tcgen05 / mma.sync micro-benchmarks next to their sources (sm_100a only).
set -e
cd "$(dirname "$0")"
for f in mma_rate umma_test; do
  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o $f $f.cu
done
