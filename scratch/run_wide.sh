#!/bin/bash
# wide bands through the headline kernel: tests touching the TMA kernel, then (13, 512, 8) timing and config 3 (fused statistics)
T=${1:-r73}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wide_bands or non_square or config2 or fused_pair or scene_windows or shapes_the_streaming or full_size or golden_degrade" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 300 python tools/stream_sweep.py "13" "8" "256,512" 4 auto 2>&1 | cut -c1-60
timeout 300 python tools/stream_sweep.py "13" "8" "512" 4 stream 2>&1 | cut -c1-60
timeout 600 python tests/run_configs.py --configs 3 --out gpurun_out/${T}_c3.json 2>/dev/null | cut -c1-420
