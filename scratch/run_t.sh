#!/bin/bash
T=${1:-r21}
shift
timeout 1200 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -40 gpurun_out/${T}_pytest.log | cut -c1-300
