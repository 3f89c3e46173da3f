#!/bin/bash
# usage: scripts/sweep_env.sh "VAR1=a VAR2=b" "VAR1=c" ...   -> prints ms/step of the default bench per setting
for cfg in "$@"; do
  ms=$(env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])")
  echo "$cfg -> $ms ms/step"
done
