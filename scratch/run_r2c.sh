#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box or tile_kernel or fused_pair" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos box,reg,stream --out gpurun_out/r2c_c5.json > gpurun_out/r2c_c5.log 2> gpurun_out/r2c_c5.err
echo "sweep exit $?" >> gpurun_out/r2c_c5.err
tail -5 gpurun_out/r2c_pytest.log
