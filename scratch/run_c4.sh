#!/bin/bash
T=${1:-r45}
timeout 600 python -m pytest tests -m gpu -x -q -k "raw_scene or scene_windows or water_mask" 2>&1 | tail -3
timeout 600 python tests/run_configs.py --configs 4 --out gpurun_out/${T}_configs.json 2>&1 | tail -1 | cut -c1-1200
