"""Tensor-level wrappers over the C ABI (allocation + argument marshalling only; all arithmetic is in
the CUDA kernels).  Layout conventions: activations NHWC bf16; conv weights bf16 [Cout][k*k][Cin]."""
import numpy as np
import torch

from . import _lib

bf16 = torch.bfloat16


def _s():
    return torch.cuda.current_stream().cuda_stream


f64 = torch.float64  # the conv epilogue's BatchNorm sums are fp64 accumulators (order independent -> reproducible)


def new_stats(c, device):
    """Zeroed [2][C] fp64 buffer for the conv epilogue's per-channel sum / sum of squares."""
    return torch.zeros((2, c), dtype=f64, device=device)


def new_scratch(c, device):
    """[2C] fp64 accumulators for the BN-backward sums (zeroed by the kernels that fill them)."""
    return torch.empty((2 * c,), dtype=f64, device=device)


class Workspace:
    """A grow-only device buffer for the split reductions (weight-gradient pixel splits, head split-K).  One per
    stream that issues such kernels: consecutive launches on a stream reuse it in order."""

    def __init__(self, zero=False):
        self.buf, self.zero = None, zero

    def get(self, nbytes, device):
        if nbytes <= 0:
            return None, 0
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            make = torch.zeros if self.zero else torch.empty
            self.buf = make((int(nbytes),), dtype=torch.uint8, device=device)
        return self.buf, self.buf.numel()


_DEFAULT_WS = {}


def _default_ws(kind, zero=False):
    """Per-(kind, stream) workspace for callers that do not manage one (tests, stand-alone op calls)."""
    key = (kind, torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    if key not in _DEFAULT_WS:
        _DEFAULT_WS[key] = Workspace(zero=zero)
    return _DEFAULT_WS[key]


def _sums(t):
    """Statistic sums as the kernels expect them: fp64 [2][C] (callers may hand in fp32 totals they computed)."""
    return t if t is None or t.dtype == f64 else t.to(f64)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PeclrKernelError("peclr_b200 kernels need CUDA tensors (there is no CPU path)")
        if t is not None and not t.is_contiguous():
            raise _lib.PeclrKernelError("peclr_b200 kernels need contiguous tensors")


# ------------------------------------------------------------------ convolutions
def conv2d_fprop(x, w, k, stride, want_stats=False, out=None, stats=None):
    """x [N,H,W,Cin] bf16, w [Cout,k*k,Cin] bf16 -> y [N,H/s,W/s,Cout] bf16 (+ stats [2,Cout] fp64: per-channel
    sum and sum of squares of the stored y)."""
    _need_cuda(x, w)
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    y = out if out is not None else torch.empty((n, h // stride, wd // stride, cout), dtype=bf16, device=x.device)
    if want_stats and stats is None:
        stats = new_stats(cout, x.device)
    _lib.call("peclr_conv2d_fprop", x, w, y, n, h, wd, cin, cout, k, stride,
              stats[0] if stats is not None else None, stats[1] if stats is not None else None, _s())
    return (y, stats) if want_stats else y


def conv2d_dgrad(dy, wt, in_shape, k, stride, out=None, accumulate=False, scatter_only=False):
    """dy [N,H/s,W/s,Cout], wt [Cin,k*k,Cout] -> dx [N,H,W,Cin].  scatter_only (1x1 / stride 2): only the sampled
    pixels of dx are written, the rest stays UNINITIALISED (for conv2d_dgrad_finish(..., acc_stride=2))."""
    _need_cuda(dy, wt)
    n, h, wd, cin = in_shape
    cout = dy.shape[-1]
    dx = out if out is not None else torch.empty(in_shape, dtype=bf16, device=dy.device)
    assert not (scatter_only and accumulate)
    _lib.call("peclr_conv2d_dgrad", dy, wt, dx, n, h, wd, cin, cout, k, stride, 2 if scatter_only else int(accumulate),
              _s())
    return dx


def conv2d_dgrad_bnreduce(dy, wt, in_shape, k, stride, bn_y, saved, gamma, beta, scratch, out=None):
    """dgrad whose epilogue also accumulates the BatchNorm-backward sums of the BN + ReLU in front of the conv
    (scratch [R][2*Cin] fp32 sets: sum g, sum g*y).  Follow with bn_backward(..., scratch=scratch, reduce_done=True)."""
    assert scratch.dtype == f64 and scratch.numel() >= 2 * in_shape[-1]
    _need_cuda(dy, wt, bn_y)
    n, h, wd, cin = in_shape
    cout = dy.shape[-1]
    dx = out if out is not None else torch.empty(in_shape, dtype=bf16, device=dy.device)
    _lib.call("peclr_conv2d_dgrad_bnreduce", dy, wt, dx, n, h, wd, cin, cout, k, stride, bn_y, saved[0], saved[1],
              gamma, beta, scratch, _s())
    return dx


def conv2d_dgrad_finish(dy, wt, dx, bn_y, mask_bits, scratch, acc_stride=1):
    """1x1 / stride-1 dgrad completing a residual block's input gradient IN PLACE: dx <- (dx + dgrad(dy)) masked with
    the previous block output's ReLU bits; scratch (fp64 [2*Cin]) <- sum g, sum g*y for that block's last BatchNorm
    (bn_y = its input).  Follow with bn_backward(dx, None, bn_y, ..., reduce_done=True).
    acc_stride=2: dx holds a gradient on its even-row / even-column pixels only (conv2d_dgrad(..., scatter_only=True)
    of a stride-2 1x1 convolution into a buffer that was never zeroed); every other pixel counts as 0."""
    _need_cuda(dy, wt, dx, bn_y, mask_bits)
    n, h, wd, cin = dx.shape
    cout = dy.shape[-1]
    assert scratch.dtype == f64 and scratch.numel() >= 2 * cin and mask_bits.dtype == torch.uint8
    assert tuple(bn_y.shape) == tuple(dx.shape) and mask_bits.numel() == n * h * wd * cin // 8
    if acc_stride == 1:
        _lib.call("peclr_conv2d_dgrad_finish", dy, wt, dx, n, h, wd, cin, cout, bn_y, mask_bits, scratch, _s())
    else:
        _lib.call("peclr_conv2d_dgrad_finish_lattice", dy, wt, dx, n, h, wd, cin, cout, bn_y, mask_bits, scratch,
                  int(acc_stride), _s())
    return dx


def conv2d_wgrad(x, dy, k, stride, dw=None, ws=None):
    """dw [Cout,k*k,Cin] fp32 += dy^T im2col(x).  ws: the Workspace the pixel splits' partials go through."""
    _need_cuda(x, dy)
    n, h, wd, cin = x.shape
    cout = dy.shape[-1]
    if dw is None:
        dw = torch.zeros((cout, k * k, cin), dtype=torch.float32, device=x.device)
    need = _lib.call("peclr_conv2d_wgrad_workspace_bytes", n, h, wd, cin, cout, k, stride)
    if need < 0:
        raise _lib.PeclrKernelError("peclr_conv2d_wgrad: unsupported geometry")
    buf, nbytes = (ws or _default_ws("wgrad")).get(need, x.device)
    _lib.call("peclr_conv2d_wgrad", x, dy, dw, n, h, wd, cin, cout, k, stride, buf, nbytes, _s())
    return dw


REDUCE_DTYPE = np.dtype([("partial", "<i8"), ("dw", "<i8"), ("n4", "<i8"), ("ksplit", "<i4"), ("blk_begin", "<i4")])


def conv2d_wgrad_partials(x, dy, k, stride, dw, ws_region):
    """Weight gradient without its reduction: the pixel splits' slabs go to `ws_region` (uint8 tensor sized by
    peclr_conv2d_wgrad_workspace_bytes; None when there is a single split, which accumulates into dw directly).  The
    caller adds the slabs to dw later, for many convolutions at once (wgrad_reduce_batched)."""
    _need_cuda(x, dy, dw)
    n, h, wd, cin = x.shape
    cout = dy.shape[-1]
    _lib.call("peclr_conv2d_wgrad_partials", x, dy, dw, n, h, wd, cin, cout, k, stride, ws_region,
              ws_region.numel() if ws_region is not None else 0, _s())


def build_reduce_table(entries, device):
    """entries: [(partial_ptr, dw_ptr, elements, ksplit)] -> (device table, rows, total blocks)."""
    per_block = _lib.call("peclr_wgrad_reduce_block_f4")
    tab = np.zeros(len(entries), dtype=REDUCE_DTYPE)
    blocks = 0
    for i, (pp, dp, n, ks) in enumerate(entries):
        assert n % 4 == 0
        tab[i] = (pp, dp, n // 4, ks, blocks)
        blocks += (n // 4 + per_block - 1) // per_block
    return torch.from_numpy(tab.view(np.uint8)).to(device), len(entries), blocks


def wgrad_reduce_batched(table):
    t, n, blocks = table
    if n:
        _lib.call("peclr_wgrad_reduce_batched", t, n, blocks, _s())


def stem_input(img1, img2, out=None):
    """two fp32 NCHW batches -> zero-padded space-to-depth bf16 [2B, H/2+3, W/2+4, 16] (2x2 pixel blocks as
    16-channel pixels, block (Y, X) at (Y+2, X+2), channel = dy*6 + dx*3 + c)."""
    _need_cuda(img1, img2)
    b, c, h, w = img1.shape
    assert c == 3 and img1.dtype == torch.float32 and (img2 is None or img2.shape == img1.shape)
    if out is None:  # img2 None: img1 alone (any batch size)
        out = torch.empty(((2 if img2 is not None else 1) * b, h // 2 + 3, w // 2 + 4, 16), dtype=bf16,
                          device=img1.device)
    _lib.call("peclr_stem_input", img1, img2, out, b, h, w, _s())
    return out


def stem_pack(w_master, out=None):
    """(64,3,7,7) fp32 channels_last master -> [64,4,64] bf16 packed stem weights (4x4 taps x 16 channels)."""
    assert w_master.shape == (64, 3, 7, 7)
    flat = w_master.permute(0, 2, 3, 1)
    assert flat.is_contiguous(), "stem weight must be stored channels_last"
    if out is None:
        out = torch.empty((64, 4, 64), dtype=bf16, device=w_master.device)
    _lib.call("peclr_stem_pack", flat, out, _s())
    return out


def stem_fprop(xpad, wpack, h, w, want_stats=False, out=None, stats=None):
    n = xpad.shape[0]
    y = out if out is not None else torch.empty((n, h // 2, w // 2, 64), dtype=bf16, device=xpad.device)
    if want_stats and stats is None:
        stats = new_stats(64, xpad.device)
    _lib.call("peclr_stem_fprop", xpad, wpack, y, n, h, w,
              stats[0] if stats is not None else None, stats[1] if stats is not None else None, _s())
    return (y, stats) if want_stats else y


def stem_wgrad(xpad, dy, h, w, dwpack=None, ws=None):
    n = xpad.shape[0]
    if dwpack is None:
        dwpack = torch.zeros((64, 4, 64), dtype=torch.float32, device=xpad.device)
    need = _lib.call("peclr_stem_wgrad_workspace_bytes", n, h, w)
    buf, nbytes = (ws or _default_ws("wgrad")).get(need, xpad.device)
    _lib.call("peclr_stem_wgrad", xpad, dy, dwpack, n, h, w, buf, nbytes, _s())
    return dwpack


# ------------------------------------------------------------------ batch norm & friends
def bn_apply(y, stats, gamma, beta, relu=True, res=None, res_bn=None, running=None, eps=1e-5, momentum=0.1, out=None,
             saved=None, rsaved=None, mask_out=None):
    """out = [relu](bn(y) + res).  res_bn = (stats, gamma, beta, running or None) if res is a raw conv output.
    Returns out, (mean, invstd)[, (rmean, rinvstd)]."""
    c = y.shape[-1]
    m = y.numel() // c
    dev = y.device
    out = out if out is not None else torch.empty_like(y)
    if saved is None:
        saved = torch.empty((2, c), dtype=torch.float32, device=dev)
    if rsaved is None and res_bn is not None:
        rsaved = torch.empty((2, c), dtype=torch.float32, device=dev)
    rstats, rgamma, rbeta, rrunning = res_bn if res_bn is not None else (None, None, None, None)
    stats, rstats = _sums(stats), _sums(rstats)
    _lib.call(
        "peclr_bn_apply", y, stats[0], stats[1], gamma, beta, res,
        rstats[0] if rstats is not None else None, rstats[1] if rstats is not None else None, rgamma, rbeta,
        out, mask_out, saved[0], saved[1],
        running[0] if running is not None else None, running[1] if running is not None else None,
        rsaved[0] if rsaved is not None else None, rsaved[1] if rsaved is not None else None,
        rrunning[0] if rrunning is not None else None, rrunning[1] if rrunning is not None else None,
        m, c, eps, momentum, int(relu), _s())
    return (out, saved, rsaved) if res_bn is not None else (out, saved)


def bn_backward(dout, mask, y, saved, gamma, dgamma, dbeta, want_g=False, scratch=None, dy=None, g_out=None,
                beta=None, reduce_done=False):
    """BatchNorm(+ReLU) backward.  ReLU mask: `mask` tensor (activation > 0) if given, else recomputed from y when
    `beta` is given, else none (dout already masked).  dgamma/dbeta are accumulated in place.  Returns dy[, g]."""
    c = y.shape[-1]
    m = y.numel() // c
    # mask: bf16 activation (mode 1) or the uint8 bit mask bn_apply wrote (mode 3)
    mode = (3 if mask.dtype == torch.uint8 else 1) if mask is not None else (2 if beta is not None else 0)
    if scratch is None:
        scratch = new_scratch(c, y.device)
    assert scratch.dtype == f64 and scratch.numel() >= 2 * c
    dy = dy if dy is not None else torch.empty_like(y)
    if want_g and g_out is None:
        g_out = torch.empty_like(y)
    if not reduce_done:  # (the sums may already come from a dgrad epilogue: conv2d_dgrad_bnreduce)
        _lib.call("peclr_bn_bwd_reduce", dout, mask, y, saved[0], saved[1], gamma, beta, mode, scratch, m, c, _s())
    _lib.call("peclr_bn_bwd_apply", dout, mask, y, saved[0], saved[1], gamma, beta, mode, scratch, dy, g_out,
              dgamma, dbeta, m, c, _s())
    return (dy, g_out) if want_g else dy


def stem_bn_relu_pool(y, stats, gamma, beta, running=None, eps=1e-5, momentum=0.1, out=None, saved=None,
                      want_idx=True):
    """Returns pooled activation, saved (mean, invstd), idx (uint8 winners, or None)."""
    n, h, w, c = y.shape
    assert c == 64
    stats = _sums(stats)
    out = out if out is not None else torch.empty((n, h // 2, w // 2, 64), dtype=bf16, device=y.device)
    idx = torch.empty((n, h // 2, w // 2, 64), dtype=torch.uint8, device=y.device) if want_idx else None
    if saved is None:
        saved = torch.empty((2, 64), dtype=torch.float32, device=y.device)
    _lib.call("peclr_stem_bn_relu_pool", y, stats[0], stats[1], gamma, beta, out, idx, saved[0], saved[1],
              running[0] if running is not None else None, running[1] if running is not None else None,
              n, h, w, eps, momentum, _s())
    return out, saved, idx


def stem_pool_bn_backward(dpool, idx, y, saved, gamma, beta, dgamma, dbeta, g_buf=None, scratch=None):
    """Backward through maxpool + relu + stem BN.  Returns dy (grad of the raw stem conv output)."""
    n, h, w, c = y.shape
    if scratch is None:
        scratch = new_scratch(64, y.device)
    assert scratch.dtype == f64 and scratch.numel() >= 128
    g = g_buf if g_buf is not None else torch.empty_like(y)
    _lib.call("peclr_stem_pool_bwd", dpool, idx, y, saved[0], saved[1], gamma, beta, g, scratch, n, h, w, _s())
    # pass 2: plain BN backward apply on the already masked gradient (in place: dy overwrites g)
    _lib.call("peclr_bn_bwd_apply", g, None, y, saved[0], saved[1], gamma, None, 0, scratch, g, None, dgamma, dbeta,
              n * h * w, 64, _s())
    return g


def avgpool_fwd(x, out=None):
    n, h, w, c = x.shape
    out = out if out is not None else torch.empty((n, c), dtype=torch.float32, device=x.device)
    _lib.call("peclr_avgpool_fwd", x, out, n, h * w, c, _s())
    return out


def avgpool_bwd(dout, shape, out=None):
    n, h, w, c = shape
    out = out if out is not None else torch.empty(shape, dtype=bf16, device=dout.device)
    _lib.call("peclr_avgpool_bwd", dout, out, n, h * w, c, _s())
    return out


# ------------------------------------------------------------------ head (fp32)
def _sgemm(a, b, c, bias, m, n, k, sam, sak, sbk, sbn, ldc, accumulate, ws):
    need = _lib.call("peclr_sgemm_workspace_bytes", m, n, k)
    buf, nbytes = (ws or _default_ws("sgemm", zero=True)).get(need, a.device)
    _lib.call("peclr_sgemm", a, b, c, bias, m, n, k, sam, sak, sbk, sbn, ldc, accumulate, buf, nbytes, _s())


def linear_fwd(x, w, bias=None, out=None, ws=None):
    """x [M,K] @ w[N,K]^T (+ bias).  ws: zero-initialised Workspace of the split-K partials (shared by the head)."""
    m, k = x.shape
    n = w.shape[0]
    out = out if out is not None else torch.empty((m, n), dtype=torch.float32, device=x.device)
    _sgemm(x, w, out, bias, m, n, k, k, 1, 1, k, n, 0, ws)
    return out


def linear_dgrad(dy, w, out=None, ws=None):
    """dy [M,N] @ w [N,K] -> [M,K]."""
    m, n = dy.shape
    k = w.shape[1]
    out = out if out is not None else torch.empty((m, k), dtype=torch.float32, device=dy.device)
    _sgemm(dy, w, out, None, m, k, n, n, 1, k, 1, k, 0, ws)
    return out


def linear_wgrad(dy, x, dw, ws=None):
    """dw [N,K] += dy[M,N]^T @ x[M,K]."""
    m, n = dy.shape
    k = x.shape[1]
    _sgemm(dy, x, dw, None, n, k, m, 1, n, k, 1, k, 1, ws)
    return dw


def bn1d_relu_fwd(x, gamma, beta, running=None, eps=1e-5, momentum=0.1):
    m, c = x.shape
    out = torch.empty_like(x)
    saved = torch.empty((2, c), dtype=torch.float32, device=x.device)
    _lib.call("peclr_bn1d_relu_fwd", x, gamma, beta, out, saved[0], saved[1],
              running[0] if running is not None else None, running[1] if running is not None else None,
              m, c, eps, momentum, _s())
    return out, saved


def bn1d_relu_bwd(dout, out, x, saved, gamma, dgamma, dbeta):
    m, c = x.shape
    dx = torch.empty_like(x)
    _lib.call("peclr_bn1d_relu_bwd", dout, out, x, saved[0], saved[1], gamma, dx, dgamma, dbeta, m, c, _s())
    return dx


def colsum_acc(x, out):
    _lib.call("peclr_colsum_acc", x, out, x.shape[0], x.shape[1], _s())
    return out


# ------------------------------------------------------------------ fused loss
STAT_NAMES = [f"proj{v}{c}_{s}" for v in (1, 2) for c in "xy" for s in ("mean", "median", "min", "max")]


def ntxent_workspace(batch, world=1, device="cuda"):
    nbytes = _lib.call("peclr_ntxent_workspace_bytes", batch, world)
    return torch.zeros((nbytes + 3) // 4, dtype=torch.float32, device=device)


def ntxent_fused(p, angle, jx, jy, image_hw, crop, rotate, temperature=0.5, want_grad=True, workspace=None,
                 world=1, rank=0, z_peers=None, flag_peers=None):
    """Returns loss [1], stats [16], g_p (or None).  p [2B,128] fp32."""
    _need_cuda(p, angle, jx, jy)
    n, d = p.shape
    b = n // 2
    dev = p.device
    if workspace is None:
        workspace = ntxent_workspace(b, world, dev)
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    stats = torch.empty((16,), dtype=torch.float32, device=dev)
    g_p = torch.empty_like(p) if want_grad else None
    _lib.call("peclr_ntxent_fused", p, angle if rotate else None, jx if crop else None, jy if crop else None,
              b, d, int(image_hw[0]), int(image_hw[1]), int(crop), int(rotate), float(temperature), loss, stats, g_p,
              workspace, workspace.numel() * 4, world, rank, z_peers, flag_peers, _s())
    return loss, stats, g_p


def ntxent_plain(z, temperature=0.5, want_grad=True, workspace=None):
    """NT-Xent on already-normalised z [2B,128].  Returns loss [1], None, g_z."""
    _need_cuda(z)
    n, d = z.shape
    if workspace is None:
        workspace = ntxent_workspace(n // 2, 1, z.device)
    loss = torch.empty((1,), dtype=torch.float32, device=z.device)
    g = torch.empty_like(z) if want_grad else None
    _lib.call("peclr_ntxent_plain", z, n // 2, d, float(temperature), loss, g, workspace, workspace.numel() * 4, _s())
    return loss, None, g


# ------------------------------------------------------------------ stand-alone equivariance operators
# (csrc/equiv_ops.cu; reference src/models/utils.py:271-364, hybrid2_model.py:92-106)
def _per_sample(v, n, dtype, device, name):
    v = torch.as_tensor(v, device=device).reshape(-1).to(dtype).contiguous()
    if v.numel() != n:
        raise ValueError("%s must hold one value per sample (%d), got %d" % (name, n, v.numel()))
    return v


def translate_encodings_(enc, tx, ty, exact=False):
    """In place on a contiguous fp32 [n][m][d] tensor."""
    _need_cuda(enc)
    assert enc.is_contiguous() and enc.dtype == torch.float32 and enc.dim() == 3
    n, m, d = enc.shape
    tx = _per_sample(tx, n, torch.float32, enc.device, "translate_x")
    ty = _per_sample(ty, n, torch.float32, enc.device, "translate_y")
    _lib.call("peclr_translate_encodings", enc, tx, ty, n, m, d, 1 if exact else 0, _s())
    return enc


def rotate_encoding_(enc, angle, rot=None):
    """In place on a contiguous fp32 [n][m][d] tensor; rot (optional fp32 [n][4]) <- {alpha, beta, off_x, off_y}."""
    _need_cuda(enc)
    assert enc.is_contiguous() and enc.dtype == torch.float32 and enc.dim() == 3
    n, m, d = enc.shape
    angle = _per_sample(angle, n, torch.float64, enc.device, "angle")
    _lib.call("peclr_rotate_encoding", enc, angle, rot if rot is not None else 0, n, m, d, _s())
    return enc


def rotate_encoding_bwd_(g, rot):
    _need_cuda(g, rot)
    assert g.is_contiguous() and g.dtype == torch.float32 and g.dim() == 3 and rot.dtype == torch.float32
    n, m, d = g.shape
    _lib.call("peclr_rotate_encoding_bwd", g, rot, n, m, d, _s())
    return g


def rotation_2d_matrix(angle, center_x, center_y, scale=1.0):
    _need_cuda(center_x)
    n = center_x.numel()
    dev = center_x.device
    angle = _per_sample(angle, n, torch.float64, dev, "angle")
    cx = _per_sample(center_x.detach(), n, torch.float32, dev, "center_x")
    cy = _per_sample(center_y.detach(), n, torch.float32, dev, "center_y")
    out = torch.empty((n, 3, 2), dtype=torch.float32, device=dev)
    _lib.call("peclr_rotation_2d_matrix", angle, cx, cy, float(scale), out, n, _s())
    return out


def projection_stats(proj):
    """fp32 [n][m][d] -> fp32 [8]: x{mean, median, min, max}, y{...} (batch means of per-sample statistics)."""
    _need_cuda(proj)
    assert proj.dtype == torch.float32 and proj.dim() == 3
    n, m, d = proj.shape
    out = torch.empty((8,), dtype=torch.float32, device=proj.device)
    _lib.call("peclr_projection_stats", proj, out, n, m, d, _s())
    return out


# ------------------------------------------------------------------ optimiser plumbing
def build_opt_tables(seg_sizes, seg_wd, device):
    """Flat-buffer segment / chunk tables for peclr_lars_adam_step."""
    chunk = _lib.call("peclr_opt_chunk_elems")
    begins = np.zeros(len(seg_sizes) + 1, dtype=np.int64)
    begins[1:] = np.cumsum(seg_sizes)
    cseg, cbeg = [], []
    for t, (b, e) in enumerate(zip(begins[:-1], begins[1:])):
        for s in range(int(b), int(e), chunk):
            cseg.append(t)
            cbeg.append(s)
    return dict(
        seg_begin=torch.tensor(begins, dtype=torch.int64, device=device),
        seg_wd=torch.tensor(np.asarray(seg_wd, dtype=np.float32), device=device),
        chunk_seg=torch.tensor(np.asarray(cseg, dtype=np.int32), device=device),
        chunk_begin=torch.tensor(np.asarray(cbeg, dtype=np.int64), device=device),
        norms=torch.zeros(2 * len(seg_sizes), dtype=torch.float64, device=device),
        num_segs=len(seg_sizes), num_chunks=len(cseg),
    )


def lars_adam_step(p, g, m, v, tables, lr, step, p_bf16=None, lars=True, betas=(0.9, 0.999), adam_eps=1e-8,
                   eta=0.02, clip=True, lars_eps=1e-8):
    _lib.call("peclr_lars_adam_step", p, g, m, v, p_bf16, tables["seg_begin"], tables["seg_wd"], tables["num_segs"],
              tables["chunk_seg"], tables["chunk_begin"], tables["num_chunks"], tables["norms"], float(lr), int(step),
              betas[0], betas[1], adam_eps, int(lars), eta, int(clip), lars_eps, _s())


TRANSPOSE_DTYPE = np.dtype([("src_off", "<i8"), ("dst_off", "<i8"), ("cout", "<i4"), ("taps", "<i4"),
                            ("cin", "<i4"), ("tile_begin", "<i4")])


def build_transpose_table(entries, device):
    """entries: list of (src_off, dst_off, cout, taps, cin).  Returns (table tensor, num_entries, total_tiles)."""
    tab = np.zeros(len(entries), dtype=TRANSPOSE_DTYPE)
    tiles = 0
    for i, (so, do, cout, taps, cin) in enumerate(entries):
        tab[i] = (so, do, cout, taps, cin, tiles)
        tiles += taps * ((cout + 31) // 32) * ((cin + 31) // 32)
    t = torch.from_numpy(tab.view(np.uint8)).to(device)
    return t, len(entries), tiles


def weight_transpose(src_flat, dst_bf16, table):
    t, n, tiles = table
    _lib.call("peclr_weight_transpose", src_flat, dst_bf16, t, n, tiles, _s())
