#!/bin/bash
# cost of the load-completion fence in the headline kernel (KMSR_TMA_NOFENCE=1 drops it)
for f in 0 1 0 1; do
KMSR_TMA_NOFENCE=$f timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 2>/dev/null | tail -1 > /tmp/fence.json
python -c "import json; d=json.load(open('/tmp/fence.json')); print('nofence', $f, d['value'], d['roofline']['frac'])"
done
