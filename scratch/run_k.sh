#!/bin/bash
T=${1:-r29}
timeout 600 python tools/bench_kernels.py --out gpurun_out/${T}_kernels.json 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
