#!/bin/bash
# timing-only variants of the tcgen05 selector convolutions (results are wrong by construction for UMMA_DBG != 0)
for d in ${@:-0 1 2}; do
  nvcc -DUMMA_DBG=$d -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/cu$d scratch/conv_umma_test.cu 2>/dev/null
  echo "== UMMA_DBG=$d"; timeout 120 /tmp/cu$d 1024 256 0 | grep -E "layer|ms |error|mismatch"
done
