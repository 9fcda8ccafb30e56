#!/bin/bash
T=${1:-r30}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/run_configs.py --configs 3,4 --c3-check 256 --out gpurun_out/${T}_configs_2gpu.json > gpurun_out/${T}_cfg2.log 2> gpurun_out/${T}_cfg2.err; echo "rc=$?"
tail -c 400 gpurun_out/${T}_cfg2.err; cut -c1-1500 gpurun_out/${T}_cfg2.log
timeout 600 python -m pytest tests -m gpu -x -q -k "dynamic_kernels or fused_pair" 2>&1 | tail -2
