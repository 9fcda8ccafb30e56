#!/bin/bash
L=${L:-1}
for d in $@; do
  nvcc -DUMMA_DBG=$d -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/c1 scratch/conv_umma_test.cu 2>/dev/null
  echo "== layer $L UMMA_DBG=$d"; timeout 120 /tmp/c1 1024 256 $L | grep -E "ms |error"
done
