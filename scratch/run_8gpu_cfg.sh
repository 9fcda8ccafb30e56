#!/bin/bash
# BASELINE configs 3 (E pair generation + statistics, 12500 patches per GPU = 100k over 8) and 4 (8k x 8k scene) on N GPUs
T=${1:-r71}
N=${2:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/run_configs.py --configs 3,4 --c3-check 256 --out gpurun_out/${T}_configs_${N}gpu.json > gpurun_out/${T}_cfg${N}.log 2> gpurun_out/${T}_cfg${N}.err; echo "rc=$?"
tail -c 300 gpurun_out/${T}_cfg${N}.err; cut -c1-1200 gpurun_out/${T}_cfg${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/${T}_bench$N.json
