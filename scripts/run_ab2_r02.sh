# 2-GPU A/B (run with gpurun --gpus 2): SMs reserved for the NCCL kernels / NCCL CTA cap
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])" $1 "$2" || tail -5 $1.err; }
run() { env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > gpurun_out/x2.json 2> gpurun_out/x2.json.err; show gpurun_out/x2.json "$1"; }
run "X=default"
run "PECLR_SM_RESERVE=8 NCCL_MAX_CTAS=8"
run "PECLR_SM_RESERVE=4 NCCL_MAX_CTAS=4"
run "PECLR_SM_RESERVE=16 NCCL_MAX_CTAS=16"
run "X=default"
run "PECLR_SM_RESERVE=8 NCCL_MAX_CTAS=8"
run "PECLR_OVERLAP_ALLREDUCE=0"
