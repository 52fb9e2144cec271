#!/bin/bash
# Builds the standalone sm_100a microbenchmarks (run them on the GPU box: ./tools/ubench/<name>).
cd "$(dirname "$0")"
for n in ex2_throughput tmem_throughput mma_throughput; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../self_forcing_b200/csrc -o $n $n.cu || exit 1
done
