#!/bin/bash
# f4 bring-up: GPU parity tests of the denoise stage, timing, ncu of nlm_kernel
mkdir -p gpurun_out
R=${1:-r57}
timeout 900 python -m pytest tests/test_gpu_denoise.py -x -q -m gpu -s > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${R}_pytest.log
timeout 600 python tests/run_denoise.py --patches 64 --out gpurun_out/${R}_denoise.json 2> gpurun_out/${R}_denoise.err; echo "run rc=$?"
tail -3 gpurun_out/${R}_denoise.err
if [ "$2" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nlm_kernel -c 1 -o gpurun_out/${R}_nlm python tests/run_denoise.py --patches 8 --steps 1 --cpu-bands 1 > gpurun_out/${R}_ncu.log 2>&1; echo "ncu rc=$?"
fi
