#!/bin/bash
# `ncu --set full` (with source) of single launches of the eager ResNet-50 step, picked by kernel-name regex and launch
# index:   ncu_pick.sh NAME:REGEX:SKIP [...]  -> gpurun_out/pick_NAME.ncu-rep  (+ pick_NAME_src.csv, the SASS page)
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --ncu-step --no-graph --warmup 3 --no-secondary --no-cpu-baseline --no-parity ${NCU_BENCH_ARGS:-}"
for a in "$@"; do
  IFS=: read n k s <<< "$a"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k --launch-skip $s --launch-count 1 -f -o $OUT/pick_$n $BENCH > $OUT/pick_$n.log 2>&1
  ncu -i $OUT/pick_$n.ncu-rep --page source --csv > $OUT/pick_${n}_src.csv 2>/dev/null
  ncu -i $OUT/pick_$n.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/pick_${n}_sum.txt
  rm -f $OUT/pick_$n.ncu-rep
  grep -m3 "gpu__time_duration\|Kernel Name" $OUT/pick_${n}_sum.txt
done
