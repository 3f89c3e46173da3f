#!/bin/bash
# Per-launch A/B of one environment switch (GPU box): two time-only launch lists of the same eager ResNet-50 step,
# joined launch by launch.  usage: ncu_ab_launches.sh VAR=a VAR=b   -> gpurun_out/ab_launches.txt
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --ncu-step --no-graph --warmup 3 --no-secondary --no-cpu-baseline --no-parity ${NCU_BENCH_ARGS:-}"
for i in 1 2; do
  v=${!i}
  env $v timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic --clock-control none --csv --log-file $OUT/ab_$i.csv $BENCH > $OUT/ab_$i.log 2>&1
done
python scripts/join_launches.py $OUT/ab_1.csv $OUT/ab_2.csv "$1" "$2" > $OUT/ab_launches.txt
tail -40 $OUT/ab_launches.txt
