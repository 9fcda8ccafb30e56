#!/bin/bash
T=${1:-r31}
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_tma -s 3 -c 1 -f -o gpurun_out/${T}_tma $CMD > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu2 rc=$?"
CMD2="python tools/stream_case.py 11 256 8"
$CMD2 > gpurun_out/${T}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_stream -s 2 -c 1 -f -o gpurun_out/${T}_stream $CMD2 > gpurun_out/${T}_ncu3.log 2>&1
echo "ncu3 rc=$?"; cat gpurun_out/${T}_plain3.log
