#!/bin/bash
T=${1:-r36}
N=${2:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "rc=$?"
cut -c1-300 gpurun_out/${T}_bench$N.json; tail -3 gpurun_out/${T}_bench$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench$N.json') if l.startswith('{')][-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
