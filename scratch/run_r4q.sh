#!/bin/bash
# final state of the session: full GPU suite, bench (1 GPU), ncu captures of the three tcgen05 launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r4q}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/cu scratch/conv_umma_test.cu 2>/dev/null
timeout 100 /tmp/cu 1024 256 0 > gpurun_out/${T}_plain.log 2>&1; grep -E "ms " gpurun_out/${T}_plain.log
for l in 1 2 3; do timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_umma --launch-skip 1 -c 1 -f -o gpurun_out/${T}_umma_l$l /tmp/cu 1024 256 $l > gpurun_out/${T}_ncu_l$l.log 2>&1; tail -1 gpurun_out/${T}_ncu_l$l.log; done
