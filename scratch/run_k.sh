#!/bin/bash
T=${1:-r29}
timeout 600 python tools/bench_kernels.py --out gpurun_out/${T}_kernels.json 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q -k "water_mask or scene_windows or band_stats or create_patches or radiance" 2>&1 | tail -3
timeout 600 python tools/run_configs.py --configs 4 --out gpurun_out/${T}_configs.json 2>&1 | tail -2 | cut -c1-900
