#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2e
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
KMSR_LIB=$PWD/kernel-modeling-super-resolution_b200/libkmsr_debug.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box or golden or config1 or config2 or wide or windows or streaming_kernel or fused" > gpurun_out/${T}_pytest_debug.log 2>&1
echo "pytest(debug lib) exit $?" >> gpurun_out/${T}_pytest_debug.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
tail -4 gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest_debug.log; tail -3 gpurun_out/${T}_bench.err; cut -c1-600 gpurun_out/${T}_bench.json
