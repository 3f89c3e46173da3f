# per-shape A/B (GPU box): conv launch tables of the current build and of peclr_b200/libpeclr_b200_old.so, joined
python -m pytest tests -m gpu -q -x -k "kernels_gpu or step_gpu" > gpurun_out/t4.log 2>&1; tail -2 gpurun_out/t4.log; grep -n "^E  \|FAILED" gpurun_out/t4.log | head
OLD=$PWD/peclr_b200/libpeclr_b200_old.so
for m in 50 152; do
  PECLR_BENCH_DUMP_LAUNCHES=gpurun_out/shapes_new_$m.txt python bench.py --model $m --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/xn_$m.json 2> gpurun_out/xn_$m.err
  PECLR_B200_LIB=$OLD PECLR_BENCH_DUMP_LAUNCHES=gpurun_out/shapes_old_$m.txt python bench.py --model $m --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/xo_$m.json 2> gpurun_out/xo_$m.err
  python - <<PY
import json
for t in ("xn_$m","xo_$m"):
    d=json.loads(open("gpurun_out/%s.json"%t).read().strip().splitlines()[-1]); print(t, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_micro_step"], d["roofline"]["wgrad_kernel"]["kernel_ms_per_micro_step"], d["clocks"]["sm_mhz"])
def load(p):
    out={}
    for l in open(p).read().splitlines()[1:]:
        f=l.split(); out[(f[0],f[1])]=(int(f[2]),float(f[3]))
    return out
a,b=load("gpurun_out/shapes_new_$m.txt"),load("gpurun_out/shapes_old_$m.txt")
tot=0
for k in a:
    if k in b:
        d=(a[k][1]-b[k][1])*a[k][0]; tot+=d
        if abs(a[k][1]-b[k][1])>0.04*b[k][1]: print("%-22s %-32s n=%d new %.1f old %.1f  total %+.0f us"%(k[0],k[1],a[k][0],a[k][1],b[k][1],d))
print("sum of deltas %+.0f us"%tot)
PY
done
