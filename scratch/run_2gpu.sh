#!/bin/bash
T=${1:-r18}
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; echo "rc=$?"
cat gpurun_out/${T}_bench2.json; tail -5 gpurun_out/${T}_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/${T}_ref2.json 2> gpurun_out/${T}_ref2.err; echo "rc=$?"
cat gpurun_out/${T}_ref2.json
