# round-2 A/B helper (GPU box): same-box timings of one environment switch, alternating runs
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], 'frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_micro_step'], d['roofline']['wgrad_kernel']['kernel_ms_per_micro_step'])" $1 "$2" || tail -5 $1.err; }
run() { env $1 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline $2 > gpurun_out/x.json 2> gpurun_out/x.json.err; show gpurun_out/x.json "$1 $2"; }
python -m pytest tests -m gpu -q -x -k "kernels_gpu or step_gpu or fullsize or models_gpu" > gpurun_out/t4.log 2>&1; tail -2 gpurun_out/t4.log; grep -n "^E  \|FAILED" gpurun_out/t4.log | head
A=${AB_A:-PECLR_FINISH_LATTICE=1}
B=${AB_B:-PECLR_FINISH_LATTICE=0}
run "$A" ""
run "$B" ""
run "$A" ""
run "$B" ""
run "$A" "--model 152 --steps 10"
run "$B" "--model 152 --steps 10"
