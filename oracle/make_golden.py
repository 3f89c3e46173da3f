"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING the reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference ships no tests or golden vectors for this path, so the
fixtures are outputs of the reference's own functions imported through
oracle/ref_shims.py:

  loss_chain.npz  vanila_contrastive_loss / rotate_encoding / translate_encodings and
                  the Hybrid2Model loss chain (hybrid2_model.py:47-90) on seeded
                  projections, fp32 and fp64, with autograd gradients and the 16
                  projection statistics (hybrid2_model.py:92-106).
  kat.npz         the K1-K3 known-answer vectors of SURVEY.md section 3.3.
  step_c1.npz     BASELINE config 1 (RN50, B=8, 64x64, seed-0 default init, structured
                  synthetic batch): loss, 16 stats, per-parameter gradient norms and a few
                  gradient slices from Hybrid2Model.training_step + backward, and the
                  parameter norms after one LARS-Adam step (the LARS/scheduler arithmetic
                  is the restated pl_bolts one -- unpinned).
  ckpt_layout.npz state_dict key order / shapes for RN50 and RN152 and the result of the
                  reference's peclr_to_torchvision round trip.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import peclr_oracle as po  # noqa: E402
from oracle.ref_shims import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_loss_chain(ref, p, angle, jx, jy, hw, crop, rotate, dtype):
    """Lines 47-90 of the reference's hybrid2_model.py driven with a given projection
    tensor (the encoder/head are bypassed; everything after them is the reference's code)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)  # rot_mat is torch.zeros(...) in the default dtype
    try:
        p = torch.tensor(p, dtype=dtype, requires_grad=True)
        b = p.shape[0] // 2

        class _Shell(ref.Hybrid2Model):
            def __init__(self):  # no encoder
                torch.nn.Module.__init__(self)
                self.config = ref.EasyDict({"augmentation": (["crop"] if crop else []) + (["rotate"] if rotate else [])})
                self.train_metrics = {}
                self.encoder = lambda x: x
                self.projection_head = lambda x: p

        shell = _Shell()
        img = torch.zeros(b, 3, hw[0], hw[1], dtype=dtype)
        batch = {
            "transformed_image1": img, "transformed_image2": img,
            "angle_1": torch.tensor(angle[:b], dtype=torch.float64), "angle_2": torch.tensor(angle[b:], dtype=torch.float64),
            "jitter_x_1": torch.tensor(jx[:b], dtype=torch.int64), "jitter_x_2": torch.tensor(jx[b:], dtype=torch.int64),
            "jitter_y_1": torch.tensor(jy[:b], dtype=torch.int64), "jitter_y_2": torch.tensor(jy[b:], dtype=torch.int64),
        }
        z1, z2 = shell.get_transformed_projections(batch)
        loss = ref.vanila_contrastive_loss(z1, z2)
        loss.backward()
        stats = {k: float(v) for k, v in shell.train_metrics.items()}
        return float(loss), p.grad.numpy().copy(), torch.cat([z1, z2]).detach().numpy().copy(), stats
    finally:
        torch.set_default_dtype(old)


def make_loss_chain(ref):
    rng = np.random.RandomState(1234)
    out = {}
    cases = []
    idx = 0
    for b, hw in ((1, (64, 64)), (2, (64, 64)), (8, (64, 64)), (32, (224, 224)), (128, (224, 224)), (5, (128, 96))):
        for crop, rotate in ((True, True), (False, False), (True, False), (False, True)):
            if b in (32, 128, 5) and (crop, rotate) in ((True, False), (False, True)):
                continue
            n = 2 * b
            kind = idx % 3
            p = rng.randn(n, 128)
            if kind == 1:  # nearly collapsed embeddings (the regime at default init)
                p = rng.randn(1, 128) + 0.05 * rng.randn(n, 128)
            if kind == 2:  # positives correlated
                p[b:] = p[:b] + 0.3 * rng.randn(b, 128)
            p = p.astype(np.float32)
            angle = np.floor(rng.uniform(-45, 45, n))
            jx = -rng.randint(0, 15, n)
            jy = -rng.randint(0, 15, n)
            name = f"c{idx}"
            l32, g32, z32, st32 = ref_loss_chain(ref, p, angle, jx, jy, hw, crop, rotate, torch.float32)
            l64, g64, z64, _ = ref_loss_chain(ref, p.astype(np.float64), angle, jx, jy, hw, crop, rotate, torch.float64)
            out.update({
                f"{name}_p": p, f"{name}_angle": angle, f"{name}_jx": jx, f"{name}_jy": jy,
                f"{name}_hw": np.array(hw), f"{name}_flags": np.array([crop, rotate]),
                f"{name}_loss32": np.float32(l32), f"{name}_loss64": np.float64(l64),
                f"{name}_g32": g32, f"{name}_g64": g64, f"{name}_z64": z64,
                f"{name}_stat_names": np.array(sorted(st32)), f"{name}_stats": np.array([st32[k] for k in sorted(st32)], dtype=np.float32),
            })
            cases.append(name)
            idx += 1
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "loss_chain.npz"), **out)
    print("loss_chain.npz:", len(cases), "cases")


def make_kat(ref):
    a = torch.arange(32, dtype=torch.float64).reshape(4, 8)
    z1 = torch.nn.functional.normalize(torch.sin(a)).float().requires_grad_()
    z2 = torch.nn.functional.normalize(torch.cos(0.5 * a)).float().requires_grad_()
    loss = ref.vanila_contrastive_loss(z1, z2)
    loss.backward()
    e = torch.nn.functional.normalize(torch.eye(4, 8))
    k1b = ref.vanila_contrastive_loss(e, e.clone())
    pts = torch.tensor([[[2.0, 1.0], [4.0, 1.0], [3.0, 4.0]]])
    k2 = ref.rotate_encoding(pts.clone(), torch.tensor([30.0], dtype=torch.float64))
    rot = ref.get_rotation_2D_matrix(torch.tensor([30.0], dtype=torch.float64), torch.tensor([3.0]), torch.tensor([2.0]), 1.0)
    k3 = ref.translate_encodings(pts.clone(), torch.tensor([0.1]), torch.tensor([-0.2]))
    k_rot90 = ref.rotate_encoding(torch.tensor([[[1.0, 0.0], [-1.0, 0.0]]]), torch.tensor([90.0], dtype=torch.float64))
    np.savez_compressed(
        os.path.join(OUT, "kat.npz"),
        k1_z1=z1.detach().numpy(), k1_z2=z2.detach().numpy(), k1_loss=np.float32(loss.item()),
        k1_dz1=z1.grad.numpy(), k1_dz2=z2.grad.numpy(), k1b_loss=np.float32(k1b.item()),
        k2_in=pts.numpy(), k2_out=k2.numpy(), k2_rot=rot.numpy(), k3_out=k3.numpy(), k_rot90=k_rot90.numpy(),
    )
    print("kat.npz: K1 loss", loss.item(), "K1b", k1b.item())


def make_step_c1(ref):
    cfg = dict(po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64))
    torch.manual_seed(0)
    model = ref.Hybrid2Model(ref.EasyDict(cfg))
    model.trainer = po._TrainerStub(world_size=1, max_epochs=100)
    batch = po.synthetic_batch(8, 64, seed=5, structured=True)
    model.train()
    out = model.training_step(batch, 0)
    out["loss"].backward()
    names = [n for n, p in model.named_parameters() if p.grad is not None]
    gnorm = np.array([float(p.grad.norm()) for n, p in model.named_parameters() if p.grad is not None])
    pnorm0 = np.array([float(p.detach().norm()) for n, p in model.named_parameters() if p.grad is not None])
    g_last = dict(model.named_parameters())["projection_head.3.weight"].grad.numpy().copy()
    g_stem = dict(model.named_parameters())["encoder.features.0.weight"].grad.numpy().copy()
    stats = {k: float(v) for k, v in out.items()}
    model.setup("fit")
    (opt,), (sch,) = model.configure_optimizers()
    # make the first step non-trivial: the schedule starts at lr 0 (warm-up from 0)
    lrs = []
    for _ in range(3):
        sch["scheduler"].step()
        lrs.append(opt.param_groups[0]["lr"])
    opt.step()
    pnorm1 = np.array([float(p.detach().norm()) for n, p in model.named_parameters() if n in set(names)])
    w_last = dict(model.named_parameters())["projection_head.3.weight"].detach().numpy().copy()
    np.savez_compressed(
        os.path.join(OUT, "step_c1.npz"),
        names=np.array(names), grad_norm=gnorm, param_norm_before=pnorm0, param_norm_after=pnorm1,
        g_head_last=g_last, g_stem=g_stem, w_head_last_after=w_last, lrs=np.array(lrs),
        metric_names=np.array(sorted(stats)), metrics=np.array([stats[k] for k in sorted(stats)]),
        n_state_dict=np.int64(len(model.state_dict())),
    )
    print("step_c1.npz: loss", stats["loss"], "params with grad", len(names), "lrs", lrs)


def make_ckpt_layout(ref):
    import torchvision

    out = {}
    for size in ("50", "152"):
        cfg = dict(po.default_config(resnet_size=size, batch_size=8, num_samples=512))
        torch.manual_seed(0)
        model = ref.Hybrid2Model(ref.EasyDict(cfg))
        sd = model.state_dict()
        out[f"rn{size}_keys"] = np.array(list(sd.keys()))
        out[f"rn{size}_shapes"] = np.array(["x".join(map(str, v.shape)) for v in sd.values()])
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "epoch=0.ckpt")
            torch.save({"state_dict": sd}, path)
            tv = getattr(torchvision.models, "resnet" + size)(weights=None)
            ref.peclr_to_torchvision(tv, path)
            tv_sd = tv.state_dict()
            feats = [(k, v) for k, v in sd.items() if "features" in k]
            ok = all(torch.equal(tv_sd[k2], v) for (k2, _), (_, v) in zip(list(tv_sd.items()), feats))
            out[f"rn{size}_roundtrip_ok"] = np.array(ok)
            out[f"rn{size}_tv_keys"] = np.array(list(tv_sd.keys())[: len(feats)])
    np.savez_compressed(os.path.join(OUT, "ckpt_layout.npz"), **out)
    print("ckpt_layout.npz:", {k: (v.shape if v.ndim else v.item()) for k, v in out.items() if "ok" in k or "keys" in k})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    make_kat(ref)
    make_loss_chain(ref)
    make_ckpt_layout(ref)
    make_step_c1(ref)


if __name__ == "__main__":
    main()
