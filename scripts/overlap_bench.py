"""Does a weight-gradient kernel (side stream) really run concurrently with an HBM-bound BatchNorm pass (main
stream)?  For a few (wgrad shape, BN shape) pairs of the ResNet-50 backward chain: device time of each kernel alone
and of both submitted to two streams, replayed from one CUDA graph (L2 evicted first).  serial = a + b; perfect
overlap = max(a, b).  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peclr_b200 import ops  # noqa: E402

N = 256
PAIRS = [  # (Cin, H, Cout, k) of the wgrad ; (M rows, C) of the BN backward pass running next to it
    ((128, 28, 128, 3), (N * 28 * 28, 512)),
    ((128, 28, 128, 3), (N * 28 * 28, 128)),
    ((256, 14, 256, 3), (N * 14 * 14, 1024)),
    ((64, 56, 64, 3), (N * 56 * 56, 256)),
    ((1024, 14, 256, 1), (N * 14 * 14, 256)),
    ((256, 56, 64, 1), (N * 56 * 56, 64)),
]


def graph_time(bodies, reps=4):
    """bodies: list of callables, body i runs on stream i (fork/join inside the capture)."""
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    main = torch.cuda.Stream(priority=-1)
    sides = [torch.cuda.Stream() for _ in bodies[1:]]

    def once():
        flush.zero_()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        for s, b in zip(sides, bodies[1:]):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                b()
        bodies[0]()
        for s in sides:
            torch.cuda.current_stream().wait_stream(s)

    with torch.cuda.stream(main):
        once()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(main):
        with torch.cuda.graph(g, stream=main):
            for _ in range(reps):
                once()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts) / reps * 1e3  # us


def main():
    t_flush = graph_time([lambda: None])
    print("flush %.1f us" % t_flush)
    print("%-26s %-18s | %8s %8s %8s | %8s %8s" % ("wgrad", "bn_bwd (mask bits, g)", "wgrad us", "bn us", "both us",
                                                 "serial", "hidden"))
    for (cin, h, cout, k), (m, c) in PAIRS:
        x = torch.randn(N, h, h, cin, device="cuda").bfloat16()
        dyc = torch.randn(N, h, h, cout, device="cuda").bfloat16()
        dw = torch.zeros(cout, k * k, cin, device="cuda")
        y = torch.randn(m, 1, 1, c, device="cuda").bfloat16()
        dout = torch.randn(m, 1, 1, c, device="cuda").bfloat16()
        dy, g = torch.empty_like(y), torch.empty_like(y)
        mask = torch.randint(0, 255, (m, c // 8), dtype=torch.uint8, device="cuda")
        gamma = torch.rand(c, device="cuda") + 0.5
        yf = y.float().view(m, c)
        saved = torch.stack([yf.mean(0), torch.rsqrt(yf.var(0, unbiased=False) + 1e-5)])
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        scratch = ops.new_scratch(c, "cuda")

        def wg():
            ops.conv2d_wgrad(x, dyc, k, 1, dw=dw)

        def bn():
            ops.bn_backward(dout, mask, y, saved, gamma, dg, db, want_g=True, scratch=scratch, dy=dy, g_out=g)

        ta = graph_time([wg]) - t_flush
        tb = graph_time([bn]) - t_flush
        tab = graph_time([bn, wg]) - t_flush
        print("%4d->%-4d %2dx%-2d k%d         %9d x %-6d | %8.1f %8.1f %8.1f | %8.1f %7.0f%%" % (
            cin, cout, h, h, k, m, c, ta, tb, tab, ta + tb, 100 * (ta + tb - tab) / max(min(ta, tb), 1e-9)))


if __name__ == "__main__":
    main()
