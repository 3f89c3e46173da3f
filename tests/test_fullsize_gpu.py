"""BASELINE.json's full sizes (2B = 256 images, 224x224): size-independent properties of the CUDA path, where the
CPU oracle would take minutes.

* adjointness of the three convolution kernels:  <conv(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(x, dy)>
  (exact in real arithmetic; bf16 output rounding is zero-mean, so the three inner products agree to ~1e-3);
* linearity of fprop in the weights;
* BatchNorm apply produces zero-mean / unit-variance channels from the epilogue's statistics;
* the full ResNet-50 step at batch 128: loss near ln(2B-1) at default init, finite statistics and gradients,
  gradient accumulation doubles the gradient, an optimiser step moves every trained tensor.

And step-level PARITY at those sizes against the oracle executed in fp32 on the GPU itself (tests/parity_util.py):
ResNet-50 B = 128 at 224 x 224 (BASELINE config 2), ResNet-152 at 224 x 224 and 64 x 64 (configs 4 / 5), and the
16-micro-batch accumulation window of config 5 on ResNet-152.
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


def _full_cfg(resnet_size, b):
    from oracle import peclr_oracle as po

    return po.default_config(resnet_size=resnet_size, batch_size=b, num_samples=b * 1000)


def test_resnet50_c2_step_parity_vs_fp32_oracle_on_gpu():
    """BASELINE config 2 at its true size (ResNet-50, B = 128 pairs, 224 x 224): loss, the 16 statistics and the
    per-group gradients of the CUDA step against the oracle run in fp32 (TF32 off) on the same GPU, same weights
    (oracle-warm-started: 40 Adam steps at B = 32) and batch; fixed tolerances of tests/parity_util.py."""
    import parity_util as pu
    from oracle import peclr_oracle as po

    b = 128
    cfg = _full_cfg("50", b)
    oracle = pu.warm_started_oracle(cfg, steps=40, batch_size=32, size=224)
    ours = pu.candidate_from(oracle, cfg)
    batch = pu.to_cuda(po.synthetic_batch(b, 224, seed=5, structured=True))
    ref, ref_g = pu.oracle_step_on_gpu(oracle, batch)
    env = pu.autocast_envelope(oracle, batch)
    got, got_g = pu.candidate_step(ours, batch)
    assert len(got) == 17
    pu.report_and_check("C2 RN50 B=128 224^2", got, got_g, ref, ref_g, envelope=env)
    # run-to-run: bit-identical (same weights, same batch)
    g1 = ours.engine.grads.clone()
    got2, _ = pu.candidate_step(ours, batch)
    assert got2["loss"] == got["loss"] and torch.equal(ours.engine.grads, g1)
    del oracle, ours
    torch.cuda.empty_cache()


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
    # a second backward of the same batch accumulates the SAME gradient once more (the forward and every reduction
    # are reproducible; only the fp32 rounding of "g + partials" vs "0 + partials" differs): direction identical to
    # 1e-6, norm doubled
    loss2 = model.training_step(batch, 0)["loss"]
    loss2.backward()
    torch.cuda.synchronize()
    assert loss2.item() == out["loss"].item()
    g2 = model.engine.grads
    ratio = float(g2.double().norm() / g1.double().norm())
    cosine = float((g2.double() @ g1.double()) / (g2.double().norm() * g1.double().norm()))
    assert abs(ratio - 2.0) < 1e-4 and cosine > 1 - 1e-6, (ratio, cosine)
    before = model.engine.flat.clone()
    for _ in range(3):
        sch["scheduler"].step()  # lr leaves the warm-up's zero
    opt.step()
    torch.cuda.synchronize()
    moved = (model.engine.flat - before).abs()
    for s in model.engine.segs:
        if not s.name.endswith("projection_head.0.bias"):  # analytically zero gradient in front of a BatchNorm
            assert float(moved[s.begin:s.begin + s.size].max()) > 0, s.name
    del model
    torch.cuda.empty_cache()


@pytest.mark.parametrize("b,size", [(16, 224), (8, 64)])
def test_resnet152_step_parity_vs_fp32_oracle_on_gpu(b, size):
    """ResNet-152 (BASELINE configs 4 / 5: 50 bottleneck blocks, 155 convolutions; the skip-path gradient is
    accumulated in bf16 by TMA reduce-add 50 times) against the fp32 oracle on the GPU at 224 x 224 (B = 16) and
    64 x 64 (B = 8), warm-started for 1000 Adam steps at B = 8 / 64 x 64 (tests/parity_util.py: why so long)."""
    import parity_util as pu
    from oracle import peclr_oracle as po

    cfg = _full_cfg("152", b)
    oracle = pu.warm_started_oracle(_full_cfg("152", 8), steps=1000, batch_size=8, size=64)
    oracle.config = cfg
    ours = pu.candidate_from(oracle, cfg)
    assert len(ours.engine.segs) == 470 and ours.engine.total == 59259456
    batch = pu.to_cuda(po.synthetic_batch(b, size, seed=5, structured=True))
    ref, ref_g = pu.oracle_step_on_gpu(oracle, batch)
    env = pu.autocast_envelope(oracle, batch)
    got, got_g = pu.candidate_step(ours, batch)
    tol = (pu.TOL_DLOSS_SMALL, pu.TOL_COS_ALL, pu.TOL_COS_TOP)
    pu.report_and_check("RN152 B=%d %d^2" % (b, size), got, got_g, ref, ref_g, envelope=env, tol=tol)
    g1 = ours.engine.grads.clone()
    got2, _ = pu.candidate_step(ours, batch)
    assert got2["loss"] == got["loss"] and torch.equal(ours.engine.grads, g1)
    del oracle, ours
    torch.cuda.empty_cache()


def test_resnet152_c5_accumulation_window():
    """BASELINE config 5's accumulation (accumulate_grad_batches 16) on ResNet-152: 16 micro-batches at scale 1/16
    through the CUDA-graph path equal the eagerly accumulated gradient bit for bit, the accumulated gradient matches
    the fp32 oracle's, and one optimiser step at the paper's lr = 1e-4 * sqrt(64 * 16) follows.  (B = 8 per micro-batch
    at 64 x 64 here; bench.py --model 152 --batch 64 --accumulate 16 runs the true size.)"""
    import parity_util as pu
    from oracle import peclr_oracle as po
    from peclr_b200.graphed import GraphedStep

    b, acc = 8, 16
    cfg = po.default_config(resnet_size="152", batch_size=64, num_samples=64 * 16 * 100, num_of_mini_batch=acc)
    oracle = pu.warm_started_oracle(_full_cfg("152", 8), steps=1000, batch_size=8, size=64)
    oracle.config = cfg
    ours = pu.candidate_from(oracle, cfg)

    class T:
        world_size, max_epochs = 1, 100

    ours.trainer = T()
    ours.setup("fit")
    (opt,), (sch,) = ours.configure_optimizers()
    assert opt.defaults["lr"] == pytest.approx(1e-4 * math.sqrt(64 * 16))
    batches = [pu.to_cuda(po.synthetic_batch(b, 64, seed=300 + i)) for i in range(acc)]
    sd = {k: v.clone() for k, v in ours.state_dict().items()}
    ours.train()
    opt.zero_grad()
    for i, bt in enumerate(batches):
        (ours.training_step(bt, i)["loss"] / acc).backward()
    torch.cuda.synchronize()
    eager = ours.engine.grads.clone()
    ours.load_state_dict(sd)
    graphed = GraphedStep(ours, batches[0], grad_scale=1.0 / acc)
    opt.zero_grad()
    for bt in batches:
        graphed(bt)
    torch.cuda.synchronize()
    assert torch.equal(ours.engine.grads, eager)
    # the accumulated gradient is the mean over the window: compare with the fp32 oracle's
    oracle.zero_grad(set_to_none=True)
    with pu.strict_fp32():
        for i, bt in enumerate(batches):
            (oracle.training_step({k: v.clone() for k, v in bt.items()}, i)["loss"] / acc).backward()
    ref_g = pu.grads_by_group(po.named_grads(oracle))
    got_g = pu.grads_by_group({n: p.grad for n, p in ours.named_parameters() if not n.startswith("encoder.final_layer")})
    cosines = {k: pu.cos(got_g[k], ref_g[k]) for k in ref_g}
    print("\n[parity C5 window RN152 16 x B=8 64^2] gradient cosines", {k: round(v, 4) for k, v in cosines.items()})
    assert cosines["all"] >= pu.TOL_COS_ALL and cosines["layer4"] >= pu.TOL_COS_TOP and cosines["head"] >= pu.TOL_COS_TOP
    before = ours.engine.flat.clone()
    for _ in range(2):
        sch["scheduler"].step()
    opt.step()
    torch.cuda.synchronize()
    assert float((ours.engine.flat - before).abs().max()) > 0
    del oracle, ours
    torch.cuda.empty_cache()
