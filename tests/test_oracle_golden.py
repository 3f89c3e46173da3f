"""The oracle restatement (oracle/peclr_oracle.py) against fixtures produced by executing
the reference (oracle/make_golden.py) -- runs anywhere, no GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import peclr_oracle as po


@pytest.fixture(scope="module")
def chain(golden_dir):
    return np.load(os.path.join(golden_dir, "loss_chain.npz"))


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "kat.npz"))


def _case(chain, name):
    g = lambda k: chain[f"{name}_{k}"]
    crop, rotate = (bool(x) for x in g("flags"))
    return dict(p=g("p"), angle=g("angle"), jitter_x=g("jx"), jitter_y=g("jy"),
                image_hw=tuple(int(x) for x in g("hw")), crop=crop, rotate=rotate)


def test_numpy_chain_fp64_matches_reference_autograd(chain):
    for name in chain["cases"]:
        kw = _case(chain, name)
        out = po.loss_chain_numpy(dtype=np.float64, **kw)
        assert abs(out["loss"] - chain[f"{name}_loss64"]) <= 1e-12 * max(1, abs(chain[f"{name}_loss64"])), name
        g64 = chain[f"{name}_g64"]
        # tolerance 1e-9 relative to the gradient scale (fp64 reference, rounding only)
        assert np.abs(out["g_p"] - g64).max() <= 1e-9 * np.abs(g64).max() + 1e-15, name  # B=1: grad is analytically 0
        assert np.abs(out["z"] - chain[f"{name}_z64"]).max() <= 1e-12, name


def test_numpy_chain_fp32_within_reference_fp32_envelope(chain):
    for name in chain["cases"]:
        kw = _case(chain, name)
        out = po.loss_chain_numpy(dtype=np.float32, **kw)
        ref32, ref64 = chain[f"{name}_loss32"], chain[f"{name}_loss64"]
        assert abs(out["loss"] - ref64) <= 2e-6 * max(1, abs(ref64)) + 2 * abs(ref32 - ref64), name
        g64 = chain[f"{name}_g64"]
        scale = np.abs(g64).max()
        err_ref = np.abs(chain[f"{name}_g32"] - g64).max()
        assert np.abs(out["g_p"] - g64).max() <= 3 * err_ref + 1e-5 * scale, name


def test_projection_stats(chain):
    for name in chain["cases"]:
        p = chain[f"{name}_p"]
        b = p.shape[0] // 2
        st = {}
        st.update(po.projection_stats_numpy(p[:b].reshape(b, -1, 2), "proj1"))
        st.update(po.projection_stats_numpy(p[b:].reshape(b, -1, 2), "proj2"))
        for k, v in zip(chain[f"{name}_stat_names"], chain[f"{name}_stats"]):
            assert abs(st[str(k)] - v) <= 2e-6 * max(1.0, abs(v)), (name, k)


def test_torch_restatement_matches_kats(kat):
    z1 = torch.tensor(kat["k1_z1"], requires_grad=True)
    z2 = torch.tensor(kat["k1_z2"], requires_grad=True)
    loss = po.vanila_contrastive_loss(z1, z2)
    loss.backward()
    assert loss.item() == pytest.approx(1.7633802891, abs=2e-7)
    assert float(kat["k1_loss"]) == pytest.approx(loss.item(), abs=1e-7)
    np.testing.assert_allclose(z1.grad.numpy(), kat["k1_dz1"], atol=1e-7)
    np.testing.assert_allclose(z2.grad.numpy(), kat["k1_dz2"], atol=1e-7)
    e = torch.nn.functional.normalize(torch.eye(4, 8))
    assert po.vanila_contrastive_loss(e, e.clone()).item() == pytest.approx(0.5944375992, abs=2e-7)  # K1b
    pts = torch.tensor(kat["k2_in"])
    out = po.rotate_encoding(pts.clone(), torch.tensor([30.0], dtype=torch.float64))
    np.testing.assert_allclose(out.numpy(), kat["k2_out"], atol=1e-7)
    np.testing.assert_allclose(out.numpy()[0, 0], [1.63397467, 1.63397455], atol=2e-7)  # K2
    out3 = po.translate_encodings(pts.clone(), torch.tensor([0.1]), torch.tensor([-0.2]))
    np.testing.assert_allclose(out3.numpy(), [[[2.2, 0.4], [4.2, 0.4], [3.2, 3.4]]], atol=1e-6)  # K3
    r90 = po.rotate_encoding(torch.tensor([[[1.0, 0.0], [-1.0, 0.0]]]), torch.tensor([90.0], dtype=torch.float64))
    np.testing.assert_allclose(r90.numpy(), kat["k_rot90"], atol=1e-7)
    np.testing.assert_allclose(r90.numpy()[0, 0], [0.0, -1.0], atol=1e-6)  # OpenCV convention KAT


def test_numpy_chain_matches_kats(kat):
    # K1 through the closed form: build p whose normalised halves are z1 / z2.
    p = np.concatenate([kat["k1_z1"], kat["k1_z2"]]).astype(np.float64)
    out = po.loss_chain_numpy(p, None, None, None, (64, 64), False, False, dtype=np.float64)
    assert out["loss"] == pytest.approx(1.76338026, abs=1e-6)


def test_schedule_recursive_equals_closed_form():
    prm = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([prm], lr=4.525e-3)
    sch = po.LinearWarmupCosineAnnealingLR(opt, warmup_epochs=40, max_epochs=400)
    for step in range(400):
        want = po.warmup_cosine_lr(step, 4.525e-3, 40, 400)
        assert opt.param_groups[0]["lr"] == pytest.approx(want, rel=1e-9, abs=1e-15), step
        sch.step()


def test_lars_adam_numpy_matches_torch_restatement():
    torch.manual_seed(3)
    for wd, lr, lars in ((1e-6, 1e-3, True), (0.0, 1e-3, True), (1e-6, 0.0, True), (1e-4, 2e-3, False)):
        p = torch.nn.Parameter(torch.randn(37, 11))
        opt = torch.optim.Adam([{"params": [p], "weight_decay": wd}], lr=lr)
        wrapped = po.LARSWrapper(opt) if lars else opt
        pn, m, v = p.detach().numpy().copy(), np.zeros((37, 11), np.float32), np.zeros((37, 11), np.float32)
        for step in range(1, 4):
            g = torch.randn(37, 11) * 10.0 ** (-step)
            p.grad = g.clone()
            wrapped.step()
            pn, m, v = po.lars_adam_step_numpy(pn, g.numpy(), m, v, step, lr, wd, lars=lars)
            np.testing.assert_allclose(pn, p.detach().numpy(), rtol=2e-6, atol=1e-7)


def test_step_c1_golden_matches_oracle_model(golden_dir):
    """BASELINE config 1: the restated model reproduces the reference's loss / grads from
    the same seed (same torch ops in the same order; thread count may differ -> 1e-3)."""
    gold = np.load(os.path.join(golden_dir, "step_c1.npz"))
    cfg = po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64)
    torch.manual_seed(0)
    model = po.OracleHybrid2Model(cfg)
    assert len(model.state_dict()) == int(gold["n_state_dict"]) == 328
    batch = po.synthetic_batch(8, 64, seed=5, structured=True)
    model.train()
    out = model.training_step(batch, 0)
    out["loss"].backward()
    metrics = dict(zip(gold["metric_names"], gold["metrics"]))
    assert out["loss"].item() == pytest.approx(metrics["loss"], abs=1e-4)
    for k, v in metrics.items():
        assert float(out[str(k)]) == pytest.approx(v, abs=1e-4 + 1e-3 * abs(v)), k
    names = [n for n, p in model.named_parameters() if p.grad is not None]
    assert names == [str(n) for n in gold["names"]]
    g = dict(model.named_parameters())["projection_head.3.weight"].grad.numpy()
    rel = np.linalg.norm(g - gold["g_head_last"]) / np.linalg.norm(gold["g_head_last"])
    assert rel < 5e-2  # default init is chaotic; identical machine gives ~0


def test_ckpt_layout_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "ckpt_layout.npz"))
    assert bool(gold["rn50_roundtrip_ok"]) and bool(gold["rn152_roundtrip_ok"])
    cfg = po.default_config(resnet_size="50", batch_size=8, num_samples=512)
    model = po.OracleHybrid2Model(cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in gold["rn50_keys"]]
    assert ["x".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in gold["rn50_shapes"]]


def test_augmentation_parameters_follow_the_reference_random_stream(golden_dir):
    """draw_view_params (host side of the GPU augmentation) re-draws, from the stored seed, exactly the parameters the
    reference's SampleAugmenter drew when the golden file was generated (angle, jitter_x/y, h, s, a, b, crop margin) --
    same use of Python's `random`, same integer truncations on the 21 joints."""
    import random

    from peclr_b200.gpu_augment import draw_view_params

    gold = np.load(os.path.join(golden_dir, "augment.npz"))
    names = ("rotate", "crop", "random_crop", "resize", "color_jitter")
    for name in (str(c) for c in gold["cases"]):
        flags = dict(zip(names, (bool(v) for v in gold[name + "_flags"])))
        rs = tuple(int(v) for v in gold[name + "_resize"])
        random.seed(int(gold[name + "_seed"]))
        for v in (1, 2):
            d = draw_view_params(gold[name + "_joints"], gold[name + "_image"].shape[:2], flags, dict(resize_shape=rs))
            got = [d["angle"], d["jitter_x"], d["jitter_y"], d["h"], d["s"], d["a"], d["b"], d["crop_margin_scale"]]
            for g, w in zip(got, gold[f"{name}_v{v}_stored"]):
                assert (g is None and np.isnan(w)) or float(g) == float(w), (name, v, got)
            box = gold[f"{name}_v{v}_box"]
            assert [d["ox"], d["oy"], d["cw"], d["ch"], d["side"]] == [int(x) for x in box], (name, v)
            if d["angle"] is not None:
                assert -45 <= d["angle"] <= 45 and float(d["angle"]).is_integer()
            assert d["jitter_x"] <= 0 or d["ox"] == 0  # = -int(U(0, 15)) unless the box was clipped at the image edge
