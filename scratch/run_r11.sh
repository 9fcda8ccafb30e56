#!/bin/bash
T=${1:-r11}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for m in 0 1 2; do
  KMSR_TMA_DEBUG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_mode$m.json 2> gpurun_out/${T}_mode$m.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_mode$m.json').read().strip().splitlines()[-1])
    print('mode $m', round(d['value']), 'pairs/s frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), d['clocks'])
except Exception as e:
    print('mode $m failed', e); print(open('gpurun_out/${T}_mode$m.err').read()[-2000:])
PY
done
