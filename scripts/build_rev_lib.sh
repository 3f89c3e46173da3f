#!/bin/bash
# usage: scripts/build_rev_lib.sh <git-rev>   -> peclr_b200/libpeclr_b200_old.so built from that revision's csrc/
# (A/B timing on one GPU box: PECLR_B200_LIB=$PWD/peclr_b200/libpeclr_b200_old.so python bench.py ...)
# Works between revisions of the SAME ABI version (the binding refuses another one); across an ABI change, export the
# whole old tree instead: mkdir ab_old && git archive <rev> | tar -x -C ab_old && (cd ab_old && python -m peclr_b200.build)
# and run bench.py in both directories (ab_old/ is git-ignored and travels with gpurun).
set -e
rev=${1:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/peclr_b200/csrc" "$tmp/include"
for f in $(git -C "$root" ls-tree --name-only "$rev" peclr_b200/csrc/); do git -C "$root" show "$rev:$f" > "$tmp/$f"; done
git -C "$root" show "$rev:include/peclr_b200.h" > "$tmp/include/peclr_b200.h"
objs=""
for src in "$tmp"/peclr_b200/csrc/*.cu; do
  fast="--use_fast_math"
  case "$(basename "$src")" in ntxent.cu|equiv_ops.cu|lars_adam.cu|head.cu|bn_act.cu|rn25d_head.cu|augment.cu) fast="";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $fast -c "$src" -o "${src%.cu}.o" &
  objs="$objs ${src%.cu}.o"
done
wait
nvcc -shared -o "$root/peclr_b200/libpeclr_b200_old.so" $objs -gencode arch=compute_100a,code=sm_100a -cudart static
rm -rf "$tmp"
echo "built $root/peclr_b200/libpeclr_b200_old.so from $rev"
