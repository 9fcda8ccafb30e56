#!/bin/bash
# register-tile kernel: parity tests, then A/B against the streaming kernel on chosen sweep cells
T=${1:-r61}; KS=${2:-"11,13,15,21,31"}; SS=${3:-"2,4"}; PS=${4:-"64,128,256,512"}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_tile" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python tools/stream_sweep.py "$KS" "$SS" "$PS" 4 stream > gpurun_out/${T}_stream.log 2>&1
timeout 600 python tools/stream_sweep.py "$KS" "$SS" "$PS" 4 reg > gpurun_out/${T}_reg.log 2>&1
paste -d'|' gpurun_out/${T}_stream.log gpurun_out/${T}_reg.log | awk -F'|' '{split($1,a," "); split($2,b," "); printf "%s %s %s stream %s reg %s x%.2f\n", a[1],a[2],a[3],a[6],b[6],a[6]/b[6]}'
