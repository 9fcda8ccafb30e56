#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/power_probe.py > gpurun_out/r2g_power.log 2>&1
cat gpurun_out/r2g_power.log
