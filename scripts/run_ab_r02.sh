# round-2 A/B helper (GPU box): tests of the changed paths, then same-box timings of the switches
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], 'frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_micro_step'], d['roofline']['wgrad_kernel']['kernel_ms_per_micro_step'])" $1 "$2" || tail -5 $1.err; }
python -m pytest tests -m gpu -q -x -k "augment or wgrad or step_gpu or stem" > gpurun_out/t3.log 2>&1; tail -3 gpurun_out/t3.log; grep -n "^E  \|FAILED\|Error" gpurun_out/t3.log | head -20
python scripts/aug_bench.py 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r50.json 2> gpurun_out/r50.json.err; show gpurun_out/r50.json "rn50 u8"
python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --e2e-input fp32 > gpurun_out/r50f.json 2> gpurun_out/r50f.json.err; show gpurun_out/r50f.json "rn50 fp32"
PECLR_OVERLAP_WGRAD=0 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --e2e-input fp32 > gpurun_out/r50n.json 2> gpurun_out/r50n.json.err; show gpurun_out/r50n.json "rn50 no-overlap-wgrad"
python bench.py --model 152 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r152.json 2> gpurun_out/r152.json.err; show gpurun_out/r152.json "rn152"
