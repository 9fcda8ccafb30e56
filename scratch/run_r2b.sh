#!/bin/bash
# ncu --set full of the box kernel on two shapes (tile mode, band mode)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/degrade_case.py 13 256 2 box 0.5 2 > gpurun_out/r2b_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_box -s 1 -c 1 -o gpurun_out/r2b_box_13_256_2 -f python tools/degrade_case.py 13 256 2 box 0.5 2 > gpurun_out/r2b_ncu1.log 2>&1
python tools/degrade_case.py 13 64 8 box 0.5 2 > gpurun_out/r2b_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_box -s 1 -c 1 -o gpurun_out/r2b_box_13_64_8 -f python tools/degrade_case.py 13 64 8 box 0.5 2 > gpurun_out/r2b_ncu2.log 2>&1
cat gpurun_out/r2b_plain1.log gpurun_out/r2b_plain2.log
