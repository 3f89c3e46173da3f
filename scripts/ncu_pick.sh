#!/bin/bash
# `ncu --set full` (with source) of single conv_gemm launches of the eager ResNet-50 step, picked by launch index:
#   ncu_pick.sh NAME:SKIP [NAME:SKIP ...]  -> gpurun_out/pick_NAME.ncu-rep
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --ncu-step --no-graph --warmup 3 --no-secondary --no-cpu-baseline --no-parity"
for a in "$@"; do
  n=${a%%:*}; s=${a##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${NCU_KERNEL:-conv_gemm} --launch-skip $s --launch-count 1 -f -o $OUT/pick_$n $BENCH > $OUT/pick_$n.log 2>&1
  ls -la $OUT/pick_$n.ncu-rep
done
