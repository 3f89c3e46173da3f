# closing captures on the final build: time + DRAM launch list (keys roofline.traffic to this build), default bench line
NCU_FULL=0 NCU_TAG=r02 bash scripts/ncu_r02.sh > gpurun_out/ncu_final.log 2>&1; tail -4 gpurun_out/ncu_final.log
python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 400 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
