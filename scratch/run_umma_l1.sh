#!/bin/bash
for d in $@; do
  nvcc -DUMMA_DBG=$d -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/c1 scratch/conv_umma_test.cu 2>/dev/null
  echo "== UMMA_DBG=$d"; timeout 120 /tmp/c1 1024 256 1 | grep -E "ms "; timeout 120 /tmp/c1 1024 256 2 | grep -E "ms "
done
