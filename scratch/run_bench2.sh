#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2k}
N=${2:-2}
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench${N}.json 2> gpurun_out/${T}_bench${N}.err
echo "bench exit $?" >> gpurun_out/${T}_bench${N}.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_rank" > gpurun_out/${T}_pytest_nccl.log 2>&1
tail -3 gpurun_out/${T}_bench${N}.err; tail -3 gpurun_out/${T}_pytest_nccl.log; cut -c1-300 gpurun_out/${T}_bench${N}.json
