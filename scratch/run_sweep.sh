#!/bin/bash
# parity of the box kernel + the config-5 sweep (usage: run_sweep.sh TAG [algos])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-sweep}
A=${2:-box,stream}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box or hand_to_each" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos $A --out gpurun_out/${T}_c5.json > gpurun_out/${T}_c5.log 2> gpurun_out/${T}_c5.err
tail -3 gpurun_out/${T}_pytest.log
