#!/bin/bash
# round 2, first GPU call: parity of the box kernel, then the config-5 sweep with every candidate kernel timed per cell
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos box,reg,stream --out gpurun_out/r2a_c5.json > gpurun_out/r2a_c5.log 2> gpurun_out/r2a_c5.err
echo "sweep exit $?" >> gpurun_out/r2a_c5.err
tail -5 gpurun_out/r2a_pytest.log
