#!/bin/bash
# ncu --set full of a list of cases: run_ncu_cases.sh TAG "K P S ALGO" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=$1; shift
for c in "$@"; do
  tag=$(echo $c | tr ' ' '_')
  python tools/degrade_case.py $c ${NCU_GB:-0.5} 2 > gpurun_out/${T}_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:degrade_ -s 1 -c 1 -o gpurun_out/${T}_$tag -f python tools/degrade_case.py $c ${NCU_GB:-0.5} 2 > gpurun_out/${T}_ncu_$tag.log 2>&1
  cat gpurun_out/${T}_plain_$tag.log
done
