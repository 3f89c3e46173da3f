#!/bin/bash
# A/B of one environment switch on ONE GPU box (boxes of the pool differ by ~4 %, so never compare across calls):
#   scripts/ab_switch.sh PECLR_WGRAD_HALO=1 ["-k wgrad or adjoint"] [--model 152]
# 1. the GPU kernel / step tests with the switch set, 2. bench with and without it, twice, interleaved.
set -u
sw=${1:?usage: ab_switch.sh VAR=value [pytest -k expression] [bench args]}
kexpr=${2:-}
shift; shift 2>/dev/null || true
cd "$(dirname "$0")/.."
if [ -n "$kexpr" ]; then sel=(-k "$kexpr"); else sel=(); fi
env "$sw" timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_step_gpu.py -m gpu -q "${sel[@]}" 2>&1 | tail -3
ms() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['final_loss'])"; }
for i in 1 2; do
  echo "with    $sw: $(env "$sw" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | ms)"
  echo "without $sw: $(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | ms)"
done
