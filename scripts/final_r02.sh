# Round-2 final validation on one B200 (run through gpurun): whole GPU suite, smoke, launch list on the final build,
# default bench line, reference arm.
python -m pytest tests -q -m gpu > gpurun_out/pytest_final_r02.log 2>&1; tail -3 gpurun_out/pytest_final_r02.log; grep -n "^E  \|FAILED" gpurun_out/pytest_final_r02.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU_FULL=1 NCU_TAG=r02 bash scripts/ncu_r02.sh > gpurun_out/ncu_final.log 2>&1; tail -12 gpurun_out/ncu_final.log
cp gpurun_out/roofline_r02.json profiles/roofline_r02.json
python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 1500 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -c 600 gpurun_out/bench_ref_r02.json
