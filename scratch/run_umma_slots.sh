#!/bin/bash
for sl in 3 4 5; do for d in 7 0; do
  nvcc -DUMMA_DBG=$d -DUMMA_SLOTS=$sl -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/cs scratch/conv_umma_test.cu 2>/dev/null
  echo "== slots $sl UMMA_DBG=$d"; timeout 120 /tmp/cs 1024 256 2 | grep -E "ms "
done; done
