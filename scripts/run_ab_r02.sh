# round-2 A/B helper (GPU box): tests of the changed paths, then same-box timings of the switches
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], 'frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_micro_step'], d['roofline']['wgrad_kernel']['kernel_ms_per_micro_step'])" $1 "$2" || tail -5 $1.err; }
python -m pytest tests -m gpu -q -x -k "augment or wgrad or step_gpu or fullsize or accumulation" > gpurun_out/t3.log 2>&1; tail -3 gpurun_out/t3.log; grep -n "^E  \|FAILED\|Error" gpurun_out/t3.log | head -20
python scripts/aug_bench.py 2>&1 | tail -4
for v in 1 0 1 0; do PECLR_BATCH_WGRAD_REDUCE=$v python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/br_$v.json 2> gpurun_out/br_$v.json.err; show gpurun_out/br_$v.json "rn50 batch_reduce=$v"; done
python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --e2e-input fp32 > gpurun_out/r50f.json 2> gpurun_out/r50f.json.err; show gpurun_out/r50f.json "rn50 fp32"
for v in 1 0; do PECLR_BATCH_WGRAD_REDUCE=$v python bench.py --model 152 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r152_$v.json 2> gpurun_out/r152_$v.json.err; show gpurun_out/r152_$v.json "rn152 batch_reduce=$v"; done
