#!/bin/bash
# round-2 evidence for the tcgen05 selector: full GPU suite, selector leg, ncu captures of the three convolution kernels
python -m pytest tests -m gpu -x -q > gpurun_out/r4c_pytest.log 2>&1; tail -3 gpurun_out/r4c_pytest.log
python tests/run_configs.py --configs 2 --out gpurun_out/r4c_selector.json 2>&1 | tail -2
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/cu scratch/conv_umma_test.cu 2>/dev/null
timeout 120 /tmp/cu 1024 256 0 > gpurun_out/r4c_plain.log 2>&1; tail -12 gpurun_out/r4c_plain.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_umma -c 3 -o gpurun_out/r4c_umma /tmp/cu 1024 256 0 > gpurun_out/r4c_ncu.log 2>&1; tail -2 gpurun_out/r4c_ncu.log
