"""CPU emulation of the index arithmetic of docs/wip/wgrad_halo.patch (NOT of the PTX semantics): per filter column s,
one zero-padded (bh + 2) x bw box of X per 64-pixel chunk; tap r = rows [r * bw, r * bw + bh * bw) of the flattened
box; accumulator g = r is written to tap index g * 3 + s.  Compared with torch's conv2d weight gradient."""
import torch


def choose(W, H):
    best = None
    for w in (8, 16, 32, 64):
        h = 64 // w
        chunks = -(-W // w) * -(-H // h)
        if best is None or chunks < best[0]:
            best = (chunks, w, h)
    return best[1], best[2]


def emulate(x, dy):  # x [N,H,W,Cin], dy [N,H,W,Cout] -> dW [Cout, 9, Cin]
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    bw, bh = choose(W, H)
    dw = torch.zeros(Cout, 9, Cin, dtype=torch.float64)

    def box(t, n0, h0, w0, hh, ww):  # TMA box with zero fill outside the tensor
        out = torch.zeros(hh, ww, t.shape[-1], dtype=torch.float64)
        for i in range(hh):
            for j in range(ww):
                h, w = h0 + i, w0 + j
                if 0 <= h < H and 0 <= w < W and n0 < N:
                    out[i, j] = t[n0, h, w]
        return out

    for s in range(3):  # tap group = filter column
        for n0 in range(N):
            for h0 in range(0, H, bh):
                for w0 in range(0, W, bw):
                    dyc = box(dy, n0, h0, w0, bh, bw).reshape(bh * bw, Cout)
                    xh = box(x, n0, h0 - 1, w0 + s - 1, bh + 2, bw).reshape((bh + 2) * bw, Cin)
                    for r in range(3):
                        rows = xh[r * bw: r * bw + bh * bw]  # descriptor start + r * bw rows, 64 rows long
                        dw[:, r * 3 + s, :] += dyc.t() @ rows
    return dw


if __name__ == "__main__":
    torch.manual_seed(0)
    for (N, H, W, Cin, Cout) in [(2, 14, 14, 4, 3), (1, 7, 7, 2, 2), (2, 28, 20, 3, 2), (1, 8, 8, 2, 2)]:
        x = torch.randn(N, H, W, Cin, dtype=torch.float64)
        dy = torch.randn(N, H, W, Cout, dtype=torch.float64)
        ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Cout, Cin, 3, 3), dy.permute(0, 3, 1, 2), padding=1)
        got = emulate(x, dy).reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
        print((N, H, W), "box", choose(W, H), "max |diff|", float((got - ref).abs().max()))
        assert torch.allclose(got, ref, atol=1e-9)
    print("index arithmetic of the wgrad halo patch: OK")
