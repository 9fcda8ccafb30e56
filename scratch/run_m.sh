#!/bin/bash
T=${1:-r38}
for m in 2 0; do
  KMSR_TMA_DEBUG=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_m$m.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_m$m.json').read().strip().splitlines()[-1]); print('${T} mode $m', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'ms', round(d['roofline']['kernel_ms'],4))
except Exception as e: print('mode $m failed', e)
PY
done
