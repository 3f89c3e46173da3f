"""Wider coverage of the module API on the B200: BasicBlock trunks, SimCLR, the op-level NT-Xent function,
gradient accumulation, the Trainer loop with checkpoints and the command-line entrypoint."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import peclr_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(resnet_size, b, head_in, augmentation=("crop", "rotate"), cls="hybrid2"):
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.simclr_model import SimCLR

    cfg = po.default_config(resnet_size=resnet_size, batch_size=b, num_samples=b * 64, augmentation=augmentation,
                            projection_head_input_dim=head_in)
    torch.manual_seed(0)
    oracle = po.OracleHybrid2Model(cfg)
    ours = (Hybrid2Model if cls == "hybrid2" else SimCLR)(EasyDict(dict(cfg)))
    ours.load_state_dict(oracle.state_dict())
    return cfg, oracle, ours.cuda()


@pytest.mark.parametrize("size,head_in,steps", [("18", 512, 60), ("34", 512, 60), ("101", 2048, 400)])
def test_other_trunks_step_parity(size, head_in, steps):
    """The other encoders the reference's -resnet_size flag selects (resnet_model.py:31-43): BasicBlock trunks
    (18 / 34: 3x3-only blocks, 64..512 channels) and ResNet-101, B = 8 at 64 x 64, oracle-warm-started, against the
    fp32 oracle on the GPU with the fixed tolerances of tests/parity_util.py."""
    import parity_util as pu

    cfg = po.default_config(resnet_size=size, batch_size=8, num_samples=8 * 64, projection_head_input_dim=head_in)
    oracle = pu.warm_started_oracle(cfg, steps=steps, batch_size=8, size=64)
    ours = pu.candidate_from(oracle, cfg)
    batch = pu.to_cuda(po.synthetic_batch(8, 64, seed=3))
    ref, ref_g = pu.oracle_step_on_gpu(oracle, batch)
    env = pu.autocast_envelope(oracle, batch)
    got, got_g = pu.candidate_step(ours, batch)
    pu.report_and_check("RN%s B=8 64^2" % size, got, got_g, ref, ref_g, envelope=env,
                        tol=(pu.TOL_DLOSS_SMALL, pu.TOL_COS_ALL, pu.TOL_COS_TOP))


def test_resnet18_default_init_loss():
    cfg, oracle, ours = _pair("18", 8, 512)
    batch = po.synthetic_batch(8, 64, seed=3)
    oracle.train(), ours.train()
    lo = oracle.training_step({k: v.clone() for k, v in batch.items()}, 0)["loss"]
    out = ours.training_step({k: v.cuda() for k, v in batch.items()}, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    # default init is chaotic for any 16-bit trunk: loss only, loose bound
    assert abs(out["loss"].item() - lo.item()) <= 0.12
    g = ours.engine.grads
    assert torch.isfinite(g).all() and float(g.norm()) > 0


def test_simclr_model_step_parity():
    """SimCLR (simclr_model.py:37-49: no equivariance correction, plain NT-Xent on the normalised projections),
    oracle-warm-started, against the fp32 oracle on the GPU."""
    import parity_util as pu
    from peclr_b200.simclr_model import SimCLR

    cfg = po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64, augmentation=())
    oracle = pu.warm_started_oracle(cfg, steps=300, batch_size=8, size=64)
    ours = pu.candidate_from(oracle, cfg, cls=SimCLR)
    batch = pu.to_cuda(po.synthetic_batch(8, 64, seed=4))
    ref, ref_g = pu.oracle_step_on_gpu(oracle, batch)
    env = pu.autocast_envelope(oracle, batch, crop=False, rotate=False)
    got, got_g = pu.candidate_step(ours, batch)
    assert set(got) == {"loss"}
    ref = {"loss": ref["loss"]}  # (the oracle class always logs the 16 statistics; SimCLR has none)
    pu.report_and_check("SimCLR RN50 B=8 64^2", got, got_g, ref, ref_g, check_stats=False, envelope=env,
                        tol=(pu.TOL_DLOSS_SMALL, pu.TOL_COS_ALL, pu.TOL_COS_TOP))


def test_encoder_accepts_any_batch_size():
    """The reference's encoder(x) takes any batch (single-image inference, a short last validation batch)."""
    import parity_util as pu

    cfg = po.default_config(resnet_size="50", batch_size=4, num_samples=4 * 64)
    oracle = pu.warm_started_oracle(cfg, steps=30, batch_size=8, size=64)
    ours = pu.candidate_from(oracle, cfg)
    ours.eval(), oracle.eval()
    g = torch.Generator().manual_seed(3)
    for n in (1, 3, 4):
        x = po.synthetic_batch(max(n, 2), 64, seed=40 + n)["transformed_image1"][:n].cuda()
        with torch.no_grad(), pu.strict_fp32():
            want = oracle.encoder(x)
            got = ours.encoder(x)
            emb = ours(x)
        assert got.shape == (n, 2048) and emb["projection"].shape == (n, 128)
        rel = float((got - want).norm() / want.norm())
        print("\n[encoder eval n=%d] rel-L2 vs fp32 oracle %.4f" % (n, rel))
        assert rel < 3e-2, (n, rel)


def test_vanila_contrastive_loss_op():
    from peclr_b200.model_utils import vanila_contrastive_loss

    g = torch.Generator().manual_seed(1)
    z = torch.nn.functional.normalize(torch.randn(64, 128, generator=g))
    z1, z2 = z[:32].clone().requires_grad_(), z[32:].clone().requires_grad_()
    ref = po.vanila_contrastive_loss(z1, z2)
    ref.backward()
    c1, c2 = z[:32].cuda().requires_grad_(), z[32:].cuda().requires_grad_()
    got = vanila_contrastive_loss(c1, c2, temperature=0.5)
    (2.0 * got).backward()
    torch.cuda.synchronize()
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.allclose(c1.grad.cpu(), 2.0 * z1.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(c2.grad.cpu(), 2.0 * z2.grad, rtol=1e-4, atol=1e-6)


def test_gradient_accumulation_and_graph_equal_eager():
    """Two micro-batches at scale 1/2 (Lightning's accumulate_grad_batches = 2) through the CUDA-graph path give
    the same accumulated gradient as the eager training_step / backward path, bit for bit, and two eager runs are
    bit-identical too."""
    from peclr_b200.graphed import GraphedStep

    cfg, oracle, ours = _pair("50", 4, 2048)
    opt = torch.optim.Adam(oracle.parameters(), lr=1e-3)
    oracle.train()
    for i in range(40):
        opt.zero_grad()
        oracle.training_step(po.synthetic_batch(4, 64, seed=200 + i), i)["loss"].backward()
        opt.step()
    ours.load_state_dict(oracle.state_dict())
    ours.train()
    batches = [{k: v.cuda() for k, v in po.synthetic_batch(4, 64, seed=10 + i).items()} for i in range(2)]
    sd = {k: v.clone() for k, v in ours.state_dict().items()}

    def eager():
        ours.load_state_dict(sd)  # (running statistics move during a pass)
        ours.zero_grad()
        for b in batches:
            (ours.training_step(b, 0)["loss"] / 2).backward()
        torch.cuda.synchronize()
        return ours.engine.grads.double().clone()

    e1, e2 = eager(), eager()
    assert torch.equal(e1, e2)  # reproducible run to run (no floating-point atomics)
    graphed = GraphedStep(ours, batches[0], grad_scale=0.5)
    ours.load_state_dict(sd)
    ours.zero_grad()
    losses = [graphed(b)["loss"].item() for b in batches]
    torch.cuda.synchronize()
    got = ours.engine.grads.double()
    assert all(np.isfinite(losses))
    assert torch.equal(got, e1)  # the captured graph submits the same kernels on the same data
    # the CUDA-graph warm-up leaves BatchNorm running statistics / num_batches_tracked where a plain run puts them
    sd_graph = {k: v.clone() for k, v in ours.state_dict().items()}
    eager()
    for k, v in ours.state_dict().items():
        if "running" in k or "num_batches" in k:
            assert torch.equal(v, sd_graph[k]), k


def test_trainer_fit_writes_reference_layout_checkpoint(tmp_path):
    from torch.utils.data import DataLoader

    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.lightning import ModelCheckpoint, Trainer
    from peclr_b200.synthetic import SyntheticTwoViewDataset

    cfg = EasyDict(dict(po.default_config(resnet_size="50", batch_size=4, num_samples=16, num_of_mini_batch=2)))
    torch.manual_seed(0)
    model = Hybrid2Model(cfg)
    data = SyntheticTwoViewDataset(16, 64, seed=5)
    loader = DataLoader(data, batch_size=4, num_workers=0, drop_last=True)
    ckpt = ModelCheckpoint(save_top_k=1, period=1, monitor="checkpoint_saving_loss", dirpath=str(tmp_path / "checkpoints"))
    trainer = Trainer(accumulate_grad_batches=2, max_epochs=2, checkpoint_callback=ckpt, default_root_dir=str(tmp_path))
    trainer.fit(model, loader, loader)
    assert trainer.global_step == 4  # 4 batches / accumulate 2, two epochs
    files = sorted(os.listdir(tmp_path / "checkpoints"))
    assert len(files) == 1 and files[0].startswith("epoch=") and files[0].endswith(".ckpt")
    sd = torch.load(tmp_path / "checkpoints" / files[0], map_location="cpu", weights_only=False)["state_dict"]
    ref_model = po.OracleHybrid2Model(po.default_config(resnet_size="50", batch_size=4, num_samples=16))
    ref_model.load_state_dict(sd)  # the reference-shaped model accepts it
    assert int(sd["encoder.features.1.num_batches_tracked"]) > 0
    assert float(model.validation_metrics_epoch["loss"]) > 0


def test_entrypoint_cli(tmp_path):
    env = dict(os.environ, SAVED_META_INFO_PATH=str(tmp_path))
    cmd = [sys.executable, os.path.join(ROOT, "src", "experiments", "peclr_training.py"), "--rotate", "--crop",
           "-resnet_size", "50", "-epochs", "1", "-batch_size", "8", "-accumulate_grad_batches", "1", "-save_top_k", "1",
           "-save_period", "1", "-num_workers", "0", "-image_size", "64", "-num_samples", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    last = json.loads(r.stdout.strip().splitlines()[-1])
    assert np.isfinite(last["epoch_loss"]) and last["images_per_sec"] > 0
    assert os.path.exists(tmp_path / "checkpoints" / "epoch=0.ckpt")
