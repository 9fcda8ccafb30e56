#!/bin/bash
# final verification of this session: smoke, full GPU suite (release + debug-assert library), both bench arms at N = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r4m}
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
KMSR_LIB=$PWD/kernel-modeling-super-resolution_b200/libkmsr_debug.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_pytest_debug.log 2>&1; echo "pytest(debug asserts) rc=$?"; tail -1 gpurun_out/${T}_pytest_debug.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
