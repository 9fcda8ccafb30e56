#!/bin/bash
# one gpurun call: GPU parity tests, then the feed / FFMA2 probes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
tail -3 gpurun_out/r10_pytest.log
{
timeout 60 ./scratch/ffma2_probe
for args in "5 8 0 0 1" "5 16 0 0 1" "5 32 0 0 1" \
            "0 4 5 8 1" "1 4 5 8 1" "2 4 5 8 1" "3 4 5 8 1" \
            "0 4 3 8 1" "2 4 3 8 1" "2 4 6 8 1" "2 8 3 8 1" "2 2 8 8 1" "2 2 4 16 1" "2 4 3 16 1" "2 2 3 32 1" "2 1 6 32 1" \
            "1 4 3 16 1" "1 2 3 32 1" "0 2 4 16 1" \
            "2 4 3 8 2" "2 2 3 16 2" "0 2 4 8 2" "1 2 4 8 2" "2 2 2 8 4" "2 1 3 16 4"; do
  timeout 60 ./scratch/feed_probe $args
done
} > gpurun_out/r10_probe.log 2>&1
cat gpurun_out/r10_probe.log
