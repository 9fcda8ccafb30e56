#!/bin/bash
# final verification of round 2: smoke, GPU tests (release + debug-assert library), both bench arms, configs 1/3/4/5,
# ncu launch list + full captures of the headline kernel and of the box kernel's cells
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2z}
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
KMSR_LIB=$PWD/kernel-modeling-super-resolution_b200/libkmsr_debug.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_pytest_debug.log 2>&1; echo "pytest(debug asserts) rc=$?"; tail -2 gpurun_out/${T}_pytest_debug.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-graph --long 0"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:degrade_tma -s 3 -c 1 -f -o gpurun_out/${T}_tma $CMD > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu2 rc=$?"
timeout 900 python tests/run_configs.py --configs 1,3,4,5 --c5-gb 4 --reps 3 --c5-algos box,stream,reg --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2> gpurun_out/${T}_configs.err; echo "configs rc=$?"
for c in "13 256 2 box" "13 256 4 box" "13 64 8 box" "31 64 8 box" "21 256 8 stream"; do
  tag=$(echo $c | awk '{print $4"_"$1"_"$2"_"$3}')
  python tools/degrade_case.py $c 0.5 2 > gpurun_out/${T}_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:degrade_ -s 1 -c 1 -o gpurun_out/${T}_$tag -f python tools/degrade_case.py $c 0.5 2 > gpurun_out/${T}_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"; cat gpurun_out/${T}_plain_$tag.log
done
