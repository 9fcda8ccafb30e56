#!/bin/bash
# the same cells through several builds of the library (usage: run_ab_libs.sh TAG "cells" "suffix suffix ...")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=$1
for sfx in $3; do
  echo "== lib$sfx"
  KMSR_LIB=$PWD/kernel-modeling-super-resolution_b200/libkmsr$sfx.so timeout 600 python tools/box_ab.py "$2" "base:" ${4:-2} ${5:-5} 2>> gpurun_out/${T}_abl.err | tee -a gpurun_out/${T}_abl.log
done
tail -3 gpurun_out/${T}_abl.err
