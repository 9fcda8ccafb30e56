#!/bin/bash
# ncu --set full of the TMA kernel in product mode and in compute-only mode (KMSR_TMA_DEBUG=2)
T=${1:-r16}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_tma -s 3 -c 1 -f -o gpurun_out/${T}_tma $CMD > gpurun_out/${T}_ncu0.log 2>&1
echo "ncu mode0 rc=$?"
KMSR_TMA_DEBUG=2 $CMD > gpurun_out/${T}_plain2.log 2>&1 &&
KMSR_TMA_DEBUG=2 ncu --set full --clock-control none --import-source on -k regex:degrade_tma -s 3 -c 1 -f -o gpurun_out/${T}_tma_mode2 $CMD > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu mode2 rc=$?"
