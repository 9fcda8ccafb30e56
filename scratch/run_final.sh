#!/bin/bash
# final verification of a round: smoke, GPU tests, both bench arms, ncu launch list + full captures
T=${1:-r55}
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_tma -s 3 -c 1 -f -o gpurun_out/${T}_tma $CMD > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu2 rc=$?"
timeout 600 python tests/run_denoise.py --patches 64 --out gpurun_out/${T}_denoise.json > /dev/null 2> gpurun_out/${T}_denoise.err; echo "denoise rc=$?"
CMD2="python tools/stream_case.py 11 256 8"
$CMD2 > gpurun_out/${T}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_stream -s 2 -c 1 -f -o gpurun_out/${T}_stream $CMD2 > gpurun_out/${T}_ncu3.log 2>&1
echo "ncu3 rc=$?"; cat gpurun_out/${T}_plain3.log
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity'])
r=json.loads([l for l in open('gpurun_out/${T}_ref.json') if l.startswith('{')][-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
CMD3="python tools/stream_sweep.py 13 2 256 2 reg"
$CMD3 > gpurun_out/${T}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_reg -s 1 -c 1 -f -o gpurun_out/${T}_reg $CMD3 > gpurun_out/${T}_ncu4.log 2>&1
echo "ncu4 rc=$?"; cat gpurun_out/${T}_plain4.log
