#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; nothing printed under ncu is a bench value).
#   1. launch list of ONE ResNet-50 step (eager submission, cudaProfilerStart/Stop around the step) with time,
#      DRAM bytes, tensor-pipe activity, L2 hit rate ... for EVERY launch -> gpurun_out/step_metrics_r02.csv
#   2. `--set full --import-source on` captures of the layer-3 convolution trio (fprop), the weight-gradient
#      kernels and the BatchNorm-backward kernels, condensed with scripts/ncu_summary.py
set -u
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__cluster_dim_x,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
BENCH="python bench.py --ncu-step --no-graph --warmup 3 --no-secondary --no-cpu-baseline ${NCU_BENCH_ARGS:-}"
TAG=${NCU_TAG:-r02}
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/step_metrics_$TAG.csv $BENCH > $OUT/ncu_list_$TAG.log 2>&1
python scripts/summarize_launches.py $OUT/step_metrics_$TAG.csv $OUT/roofline_$TAG.json > $OUT/launches_${TAG}_summary.txt 2>&1
if [ "${NCU_FULL:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm --launch-skip 24 --launch-count 8 -f -o $OUT/conv_l3_$TAG $BENCH > $OUT/ncu_full1_$TAG.log 2>&1
  ncu -i $OUT/conv_l3_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/conv_gemm_layer3_${TAG}_ncu_full.txt
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_wgrad --launch-skip 0 --launch-count 10 -f -o $OUT/wgrad_$TAG $BENCH > $OUT/ncu_full2_$TAG.log 2>&1
  ncu -i $OUT/wgrad_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/conv_wgrad_${TAG}_ncu_full.txt
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bn_bwd --launch-skip 0 --launch-count 8 -f -o $OUT/bnbwd_$TAG $BENCH > $OUT/ncu_full3_$TAG.log 2>&1
  ncu -i $OUT/bnbwd_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/bn_bwd_${TAG}_ncu_full.txt
  ls -la $OUT/*.ncu-rep
fi
cat $OUT/launches_${TAG}_summary.txt
