# Round-2 closing run on one B200 (through gpurun): whole GPU suite, smoke, default bench line (the ncu captures of
# scripts/final_r02.sh stay valid while the kernel sources -- the build digest -- are unchanged)
python -m pytest tests -q -m gpu > gpurun_out/pytest_final_r02.log 2>&1; tail -3 gpurun_out/pytest_final_r02.log; grep -n "^E  \|FAILED" gpurun_out/pytest_final_r02.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 600 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
