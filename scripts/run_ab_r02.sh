# round-2 A/B helper (GPU box): same-box timings of switches
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], 'frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_micro_step'], d['roofline']['wgrad_kernel']['kernel_ms_per_micro_step'])" $1 "$2" || tail -5 $1.err; }
run() { env $1 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/x.json 2> gpurun_out/x.json.err; show gpurun_out/x.json "rn50 $1"; }
run "PECLR_PDL=0"
run "PECLR_PDL=1 PECLR_OVERLAP_WGRAD=0"
run "PECLR_ELT_ROWS=2"
run "PECLR_ELT_WAVES=3"
run "PECLR_GRAPH_PRIORITY=0"
run "PECLR_PDL=0"
