"""Achieved HBM GB/s of the elementwise / reduction kernels on the trunk's activation shapes (2B = 256)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peclr_b200 import ops  # noqa: E402

SHAPES = [(256 * 56 * 56, 64), (256 * 56 * 56, 256), (256 * 28 * 28, 128), (256 * 28 * 28, 512), (256 * 14 * 14, 256),
          (256 * 14 * 14, 1024), (256 * 7 * 7, 512), (256 * 7 * 7, 2048)]


def timeit(fn, reps=5):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    print("waves", os.environ.get("PECLR_ELT_WAVES", "default"))
    print("%-18s | %-22s | %-22s | %-22s | %-22s" % ("M x C", "bn_apply ms GB/s", "bn_apply+res", "bwd(mode2) ms GB/s", "bwd(mask,g) ms GB/s"))
    for m, c in SHAPES:
        y = torch.randn(m, 1, 1, c, device="cuda").bfloat16()
        res = torch.randn(m, 1, 1, c, device="cuda").bfloat16()
        dout = torch.randn(m, 1, 1, c, device="cuda").bfloat16()
        out = torch.empty_like(y)
        dy = torch.empty_like(y)
        g = torch.empty_like(y)
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        yf = y.float().view(m, c)
        stats = torch.stack([yf.sum(0), (yf * yf).sum(0)])
        if hasattr(ops, "_sums"):
            stats = ops._sums(stats)  # the kernels' own [2][R][C] fp64 layout (no conversion inside the timed call)
        saved = torch.stack([yf.mean(0), torch.rsqrt(yf.var(0, unbiased=False) + 1e-5)])
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        scratch = ops.new_scratch(c, "cuda")
        nb = m * c * 2 / 1e9
        t1 = timeit(lambda: ops.bn_apply(y, stats, gamma, beta, out=out, saved=saved.clone()))
        t2 = timeit(lambda: ops.bn_apply(y, stats, gamma, beta, res=res, out=out, saved=saved.clone()))
        t3 = timeit(lambda: ops.bn_backward(dout, None, y, saved, gamma, dg, db, scratch=scratch, dy=dy, beta=beta))
        t4 = timeit(lambda: ops.bn_backward(dout, out, y, saved, gamma, dg, db, want_g=True, scratch=scratch, dy=dy, g_out=g))
        print("%9d x %-6d | %7.3f %8.0f       | %7.3f %8.0f       | %7.3f %8.0f       | %7.3f %8.0f" % (
            m, c, t1, 2 * nb / t1 * 1e3, t2, 3 * nb / t2 * 1e3, t3, 5 * nb / t3 * 1e3, t4, 8 * nb / t4 * 1e3))




def stem_main():
    """Stem pooling kernels at full size (2B = 256, 112 x 112 x 64 conv output): ms and algorithmic GB/s."""
    n, hw = 256, 112
    y = torch.randn(n, hw, hw, 64, device="cuda").bfloat16()
    gamma, beta = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.3
    yf = y.float().reshape(-1, 64)
    stats = torch.stack([yf.sum(0), (yf * yf).sum(0)])
    if hasattr(ops, "_sums"):
        stats = ops._sums(stats)
    out, saved, idx = ops.stem_bn_relu_pool(y, stats, gamma, beta)
    dpool = torch.randn_like(out)
    dg, db = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    scratch = ops.new_scratch(64, "cuda")
    g_buf = torch.empty_like(y)
    t_f = timeit(lambda: ops.stem_bn_relu_pool(y, stats, gamma, beta, out=out, saved=saved))
    t_b = timeit(lambda: _lib_call_pool_bwd(dpool, idx, y, saved, gamma, beta, g_buf, scratch))
    ye, oe = y.numel(), out.numel()
    print("stem_bn_relu_pool  %.3f ms  %6.0f GB/s" % (t_f, (2 * ye + 3 * oe) / t_f / 1e6))
    print("stem_pool_bwd      %.3f ms  %6.0f GB/s" % (t_b, (4 * ye + 3 * oe) / t_b / 1e6))


def _lib_call_pool_bwd(dpool, idx, y, saved, gamma, beta, g_buf, scratch):
    from peclr_b200 import _lib

    n, h, w, _ = y.shape
    _lib.call("peclr_stem_pool_bwd", dpool, idx, y, saved[0], saved[1], gamma, beta, g_buf, scratch, n, h, w,
              _lib.stream_ptr())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stem":
        stem_main()
    else:
        main()
