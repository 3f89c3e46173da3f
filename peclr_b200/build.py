"""Builds libpeclr_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpeclr_b200.so")
SOURCES = ["abi.cu", "conv_tc.cu", "conv_ops.cu", "bn_act.cu", "head.cu", "ntxent.cu", "equiv_ops.cu", "lars_adam.cu",
           "rn25d_head.cu", "augment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
]
# --use_fast_math would also flush denormals / relax the fp32 loss chain: those files opt out
PRECISE = {"ntxent.cu", "equiv_ops.cu", "lars_adam.cu", "head.cu", "bn_act.cu", "rn25d_head.cu", "augment.cu"}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if not name.endswith((".cu", ".cuh", ".h")):  # sources only: objects and the stamp itself are outputs
                continue
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode() + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(HERE, "csrc", ".build_stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        flags = [f for f in NVCC_FLAGS if not (src in PRECISE and f == "--use_fast_math")]
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
        if "spill" in out and any(" 0 bytes spill" not in l for l in out.splitlines() if "spill" in l):
            print(f"[peclr_b200.build] warning: register spills in {src}", file=sys.stderr)
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
