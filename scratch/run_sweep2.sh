#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-sweep}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box or hand_to_each" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos box --out gpurun_out/${T}_c5.json > gpurun_out/${T}_c5.log 2> gpurun_out/${T}_c5.err
KMSR_BOX_PW=16 timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos box --out gpurun_out/${T}_c5_pw16.json > gpurun_out/${T}_c5_pw16.log 2> gpurun_out/${T}_c5_pw16.err
tail -3 gpurun_out/${T}_pytest.log
