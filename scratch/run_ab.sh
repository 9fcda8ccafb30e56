#!/bin/bash
# A/B of box-kernel variants (usage: run_ab.sh TAG "cells" "variants")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab}
timeout 900 python tools/box_ab.py "$2" "$3" ${4:-2} ${5:-5} > gpurun_out/${T}_ab.log 2> gpurun_out/${T}_ab.err
echo "exit $?" >> gpurun_out/${T}_ab.err
cat gpurun_out/${T}_ab.log; tail -5 gpurun_out/${T}_ab.err
