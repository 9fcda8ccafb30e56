#!/bin/bash
# after the tcgen05 selector: smoke, debug-assert library on the parity suite, both bench arms at N = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r4f}
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log | cut -c1-200
KMSR_LIB=$PWD/kernel-modeling-super-resolution_b200/libkmsr_debug.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_pytest_debug.log 2>&1; echo "pytest(debug asserts) rc=$?"; tail -2 gpurun_out/${T}_pytest_debug.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
