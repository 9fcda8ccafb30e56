#!/bin/bash
# streaming kernel: its parity tests, then timing of the factor-4 / factor-8 sweep cells
T=${1:-r76}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic_streaming or non_square or shapes_the_streaming or error_codes or new_entry or interior_path" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python tools/stream_sweep.py "11,13,15,21,31" "4,8" "64,128,256,512" 4 stream 2>&1 | cut -c1-48 > gpurun_out/${T}_stream.log
cat gpurun_out/${T}_stream.log
