"""BASELINE.json's full sizes (2B = 256 images, 224x224): size-independent properties of the CUDA path, where the
CPU oracle would take minutes.

* adjointness of the three convolution kernels:  <conv(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(x, dy)>
  (exact in real arithmetic; bf16 output rounding is zero-mean, so the three inner products agree to ~1e-3);
* linearity of fprop in the weights;
* BatchNorm apply produces zero-mean / unit-variance channels from the epilogue's statistics;
* the full ResNet-50 step at batch 128: loss near ln(2B-1) at default init, finite statistics and gradients,
  gradient accumulation doubles the gradient, an optimiser step moves every trained tensor.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

N = 256
SHAPES = [  # Cin, H, Cout, k, stride  (largest-traffic and dominant-FLOP shapes of SURVEY table A2)
    (64, 56, 256, 1, 1), (256, 56, 64, 1, 1), (64, 56, 64, 3, 1), (128, 56, 128, 3, 2), (256, 56, 512, 1, 2),
    (256, 14, 1024, 1, 1), (1024, 14, 256, 1, 1), (256, 14, 256, 3, 1), (512, 7, 512, 3, 1), (2048, 7, 512, 1, 1),
]


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_adjointness_full_size(shape):
    from peclr_b200 import ops

    cin, h, cout, k, s = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, h, h, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, k * k, cin, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16()
    dy = torch.randn(N, h // s, h // s, cout, device="cuda", generator=g).bfloat16()
    y, stats = ops.conv2d_fprop(x, w, k, s, want_stats=True)
    dx = ops.conv2d_dgrad(dy, w.permute(2, 1, 0).contiguous(), tuple(x.shape), k, s)
    dw = ops.conv2d_wgrad(x, dy, k, s)
    torch.cuda.synchronize()
    s1, s2, s3 = _dot(y, dy), _dot(x, dx), _dot(w, dw)
    scale = math.sqrt(_dot(y, y) * _dot(dy, dy))
    assert abs(s1 - s2) <= 2e-3 * scale and abs(s1 - s3) <= 2e-3 * scale, (s1, s2, s3, scale)
    # the epilogue statistics are the sums of what was stored
    yf = y.float().reshape(-1, cout)
    tot = stats.float()
    assert torch.allclose(tot[0], yf.sum(0), rtol=2e-3, atol=2e-3 * float(yf.abs().sum(0).max()))
    assert torch.allclose(tot[1], (yf * yf).sum(0), rtol=2e-3)
    # linearity in the weights: conv(x, 2w) = 2 conv(x, w) exactly (power-of-two scaling commutes with rounding)
    y2 = ops.conv2d_fprop(x, (w.float() * 2).bfloat16(), k, s)
    assert torch.equal(y2.float(), 2 * y.float())


def test_bn_apply_normalises_full_size():
    from peclr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(N, 56, 56, 64, device="cuda", generator=g).bfloat16()
    w = (torch.randn(256, 1, 64, device="cuda", generator=g) / 8).bfloat16()
    y, stats = ops.conv2d_fprop(x, w, 1, 1, want_stats=True)
    ones, zeros = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    out, saved = ops.bn_apply(y, stats, ones, zeros, relu=False)
    torch.cuda.synchronize()
    of = out.float().reshape(-1, 256)
    assert float(of.mean(0).abs().max()) < 5e-3 and float((of.var(0, unbiased=False) - 1).abs().max()) < 1e-2


def test_resnet50_full_batch_step_properties():
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.synthetic import synthetic_batch

    b = 128
    cfg = EasyDict(batch_size=b, lr=1e-4, opt_weight_decay=1e-6, output_dim=128, projection_head_hidden_dim=512,
                   projection_head_input_dim=2048, warmup_epochs=10, num_of_mini_batch=1,
                   augmentation=["crop", "rotate"], optimizer="LARS", resnet_size="50", num_samples=b * 1000)
    torch.manual_seed(0)
    model = Hybrid2Model(cfg).cuda()

    class T:
        world_size, max_epochs = 1, 100

    model.trainer = T()
    model.setup("fit")
    (opt,), (sch,) = model.configure_optimizers()
    batch = {k: v.cuda() for k, v in synthetic_batch(b, 224, seed=5, structured=True).items()}
    model.train()
    opt.zero_grad()
    out = model.training_step(batch, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    assert len(out) == 17 and all(torch.isfinite(v).all() for v in out.values())
    assert abs(out["loss"].item() - math.log(2 * b - 1)) < 0.5  # embeddings nearly collapsed at default init
    g1 = model.engine.grads.clone()
    assert torch.isfinite(g1).all() and float(g1.norm()) > 0
    # every trained tensor received a gradient
    for s in model.engine.segs:
        assert float(g1[s.begin:s.begin + s.size].abs().max()) > 0 or s.name.endswith("projection_head.0.bias"), s.name
    # a second backward of the same batch accumulates: direction unchanged, norm about doubled (BN statistics of
    # the same batch are identical; only atomics' summation order differs)
    model.training_step(batch, 0)["loss"].backward()
    torch.cuda.synchronize()
    g2 = model.engine.grads
    ratio = float(g2.norm() / g1.norm())
    cosine = float((g2.double() @ g1.double()) / (g2.double().norm() * g1.double().norm()))
    assert 1.5 < ratio < 2.5 and cosine > 0.7, (ratio, cosine)
    before = model.engine.flat.clone()
    for _ in range(3):
        sch["scheduler"].step()  # lr leaves the warm-up's zero
    opt.step()
    torch.cuda.synchronize()
    moved = (model.engine.flat - before).abs()
    for s in model.engine.segs:
        if not s.name.endswith("projection_head.0.bias"):  # analytically zero gradient in front of a BatchNorm
            assert float(moved[s.begin:s.begin + s.size].max()) > 0, s.name
