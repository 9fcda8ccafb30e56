#!/bin/bash
# streaming-kernel tests + the config 5 sweep
T=${1:-r27}
timeout 1200 python -m pytest tests -m gpu -x -q -k "generic_streaming or golden_degrade or non_square or hand_to_each" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
python tests/run_configs.py --configs 5 --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2> gpurun_out/${T}_configs.err; echo rc=$?; tail -c 600 gpurun_out/${T}_configs.err
