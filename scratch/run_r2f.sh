#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2f}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box or hand_to_each" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python tests/run_configs.py --configs 5 --c5-gb 2 --reps 3 --c5-algos box,stream --out gpurun_out/${T}_c5.json > gpurun_out/${T}_c5.log 2> gpurun_out/${T}_c5.err
for c in "13 256 2" "13 64 8"; do
  tag=$(echo $c | tr ' ' '_')
  python tools/degrade_case.py $c box 0.5 2 > gpurun_out/${T}_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:degrade_box -s 1 -c 1 -o gpurun_out/${T}_box_$tag -f python tools/degrade_case.py $c box 0.5 2 > gpurun_out/${T}_ncu_$tag.log 2>&1
done
tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_plain_*.log
