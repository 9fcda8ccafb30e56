#!/bin/bash
T=${1:-r19}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 1500 python tests/run_configs.py --configs ${2:-1,3,4,5} --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2> gpurun_out/${T}_configs.err; echo "configs rc=$?"
tail -c 1500 gpurun_out/${T}_configs.err
cut -c1-1200 gpurun_out/${T}_configs.log
