"""Step-level parity on the B200 (BASELINE config 1: RN50, B=8, 64x64, synthetic two-view): the CUDA step
(Hybrid2Model.training_step + loss.backward() + optimizer step through the public module API) against the
CPU oracle on identical weights and batch.

Tolerances (SURVEY.md 3.6 / 8(d)(iii), bf16 tensor-core trunk, errors vs the fp32 oracle from an
oracle-warm-started state), FIXED: global gradient cosine >= 0.93, layer4 + head cosine >= 0.985, |dloss| <= 1e-2
at this 16-row configuration (1.5e-3 at BASELINE's B = 128: tests/parity_util.py says why).  What the reference itself does under torch bf16 autocast on the same weights / batch
is measured alongside and printed for context; it does not enter any assertion.  At default init only the loss is
compared (the gradients of ANY 16-bit trunk are uncorrelated with fp64 there).  The step is bit-reproducible: two
runs give identical loss, gradients and updated weights (no floating-point atomics anywhere, ABI 3).
"""
import copy
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import peclr_oracle as po
from parity_util import TOL_COS_ALL, TOL_COS_TOP, TOL_DLOSS_SMALL, cos, grads_by_group

pytestmark = pytest.mark.gpu

B, SIZE = 8, 64


def make_pair(warm_steps):
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    cfg = po.default_config(resnet_size="50", batch_size=B, num_samples=B * 64)
    torch.manual_seed(0)
    oracle = po.OracleHybrid2Model(cfg)
    if warm_steps:
        # the warm-up only produces WEIGHTS (Adam steps of the oracle, run on the GPU because 300 CPU steps take
        # minutes); the reference values below come from the CPU oracle, bit-identical to the executed reference
        import parity_util as pu

        oracle.load_state_dict({k: v.cpu() for k, v in
                                pu.warm_started_oracle(cfg, warm_steps, B, SIZE).state_dict().items()})
    oracle.train_metrics, oracle.plot_params = {}, {}  # hold graph tensors: not deep-copyable
    torch.manual_seed(0)
    ours = Hybrid2Model(EasyDict(dict(cfg)))
    ours.load_state_dict(oracle.state_dict())
    ours.cuda()
    return cfg, oracle, ours


@pytest.fixture(scope="module")
def warm():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return make_pair(warm_steps=300)


def test_step_loss_stats_grads_vs_oracle(warm):
    cfg, oracle, ours = warm
    batch = po.synthetic_batch(B, SIZE, seed=5)
    oracle.train()
    o = copy.deepcopy(oracle)
    o.zero_grad()
    out_o = o.training_step({k: v.clone() for k, v in batch.items()}, 0)
    out_o["loss"].backward()
    ref = grads_by_group(po.named_grads(o))
    # the reference's own envelope under bf16 autocast of trunk + head (fp32 loss chain), same weights/batch
    o2 = copy.deepcopy(oracle)
    o2.zero_grad()
    x = torch.cat((batch["transformed_image1"], batch["transformed_image2"]))
    with torch.autocast("cpu", dtype=torch.bfloat16):
        proj = o2.projection_head(o2.encoder(x))
    proj = proj.float()
    chain = po.loss_chain_numpy(proj.detach().numpy(), torch.cat((batch["angle_1"], batch["angle_2"])).numpy(),
                                torch.cat((batch["jitter_x_1"], batch["jitter_x_2"])).numpy(),
                                torch.cat((batch["jitter_y_1"], batch["jitter_y_2"])).numpy(), (SIZE, SIZE), True, True,
                                dtype=np.float32)
    proj.backward(torch.tensor(chain["g_p"]))
    env = grads_by_group(po.named_grads(o2))
    env_dloss = abs(float(chain["loss"]) - out_o["loss"].item())
    env_cos = {k: cos(env[k], ref[k]) for k in ref}

    ours.train()
    ours.zero_grad()
    ours.engine.zero_grad()
    gb = {k: v.cuda() for k, v in batch.items()}
    out = ours.training_step(gb, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    got = grads_by_group({n: p.grad for n, p in ours.named_parameters() if not n.startswith("encoder.final_layer")})
    dloss = abs(out["loss"].item() - out_o["loss"].item())
    got_cos = {k: cos(got[k], ref[k]) for k in ref}
    print("\n[parity C1 warm] loss ours %.6f oracle %.6f |d| %.2e (autocast-reference |d| %.2e)" %
          (out["loss"].item(), out_o["loss"].item(), dloss, env_dloss))
    for k in ("stem", "layer1", "layer2", "layer3", "layer4", "head", "all"):
        rel = float((got[k] - ref[k]).norm() / ref[k].norm())
        print("  %-7s |g| %.3e  cos ours %.4f  autocast-reference %.4f  rel-L2 ours %.3f" %
              (k, float(ref[k].norm()), got_cos[k], env_cos[k], rel))
    assert set(out) == set(out_o) and len(out) == 17
    assert dloss <= TOL_DLOSS_SMALL, dloss
    for k in out_o:
        if k != "loss":
            assert abs(out[k].item() - out_o[k].item()) <= 3e-2 * (abs(out_o[k].item()) + 0.05), k
    assert got_cos["all"] >= TOL_COS_ALL, got_cos
    assert got_cos["layer4"] >= TOL_COS_TOP and got_cos["head"] >= TOL_COS_TOP, got_cos


def test_step_is_bit_reproducible(warm):
    """Same weights, same batch, three runs (eager, eager, CUDA-graph replay): identical loss, statistics, gradients;
    and the optimiser step that follows gives identical weights."""
    from peclr_b200.graphed import GraphedStep

    cfg, oracle, ours = warm
    sd = {k: v.clone() for k, v in ours.state_dict().items()}
    batch = {k: v.cuda() for k, v in po.synthetic_batch(B, SIZE, seed=5).items()}

    class T:
        world_size, max_epochs = 1, 100

    def run(graphed=None):
        ours.load_state_dict(sd)
        ours.trainer = T()
        ours.setup("fit")
        (opt,), (sch,) = ours.configure_optimizers()
        ours.engine.exp_avg = ours.engine.exp_avg_sq = None
        for _ in range(3):
            sch["scheduler"].step()
        ours.train()
        opt.zero_grad()
        if graphed is None:
            out = ours.training_step(batch, 0)
            out["loss"].backward()
        else:
            out = graphed(batch)
        torch.cuda.synchronize()
        grads = ours.engine.grads.clone()
        opt.step()
        torch.cuda.synchronize()
        return {k: v.detach().clone() for k, v in out.items()}, grads, ours.engine.flat.clone()

    o1, g1, w1 = run()
    o2, g2, w2 = run()
    assert all(torch.equal(o1[k], o2[k]) for k in o1), {k: (float(o1[k]), float(o2[k])) for k in o1}
    assert torch.equal(g1, g2) and torch.equal(w1, w2)
    assert float(g1.abs().max()) > 0 and not torch.equal(w1, ours.engine.flat * 0)
    ours.load_state_dict(sd)
    graphed = GraphedStep(ours, batch, grad_scale=1.0)
    o3, g3, w3 = run(graphed)
    assert all(torch.equal(o1[k], o3[k]) for k in o1)
    assert torch.equal(g1, g3) and torch.equal(w1, w3)
    ours.load_state_dict(sd)


def test_optimizer_step_and_schedule_vs_oracle(warm):
    """Same gradients fed to the reference optimiser stack (LARSWrapper(Adam) + warm-up cosine) and to the fused
    CUDA step; the parameters must agree after each of three steps."""
    cfg, oracle, ours = warm
    o = copy.deepcopy(oracle)
    o.trainer = po._TrainerStub(world_size=1, max_epochs=100)
    o.setup("fit")
    (opt_o,), (sch_o,) = o.configure_optimizers()

    class T:
        world_size, max_epochs = 1, 100

    ours.trainer = T()
    ours.setup("fit")
    (opt,), (sch,) = ours.configure_optimizers()
    assert [len(g["params"]) for g in opt.param_groups] == [len(g["params"]) for g in opt_o.param_groups]
    batch = {k: v.cuda() for k, v in po.synthetic_batch(B, SIZE, seed=6).items()}
    before = ours.state_dict()["projection_head.3.weight"].clone()
    for step in range(3):
        sch["scheduler"].step(), sch_o["scheduler"].step()
        assert opt.param_groups[0]["lr"] == pytest.approx(opt_o.param_groups[0]["lr"], rel=1e-12)
        opt.zero_grad()
        ours.training_step(batch, step)["loss"].backward()
        mine = {n: p for n, p in ours.named_parameters()}
        for n, p in o.named_parameters():
            p.grad = None if n.startswith("encoder.final_layer") else mine[n].grad.detach().cpu().clone()
        opt.step(), opt_o.step()
        torch.cuda.synchronize()
        worst = 0.0
        for n, p in o.named_parameters():
            q = mine[n].detach().cpu()
            worst = max(worst, float((q - p.detach()).abs().max() / p.detach().abs().max().clamp_min(1e-12)))
        assert worst <= 5e-6, (step, worst)
    assert not torch.equal(before, ours.state_dict()["projection_head.3.weight"])
    # the bf16 operand copies follow the master weights
    eng = ours.engine
    assert torch.equal(eng.w_bf16, eng.flat.bfloat16())


def test_default_init_loss_only():
    cfg, oracle, ours = make_pair(warm_steps=0)
    batch = po.synthetic_batch(B, SIZE, seed=5)
    oracle.train(), ours.train()
    lo = oracle.training_step({k: v.clone() for k, v in batch.items()}, 0)["loss"].item()
    lg = ours.training_step({k: v.cuda() for k, v in batch.items()}, 0)["loss"].item()
    print("\n[parity C1 default init] loss ours %.5f oracle %.5f ln(2B-1)=%.5f" % (lg, lo, np.log(2 * B - 1)))
    assert abs(lg - lo) <= 0.12  # reference under bf16 autocast moves by 2.7e-2 here; run-to-run 5e-2


def test_running_stats_and_eval_mode(warm):
    cfg, oracle, ours = warm
    o = copy.deepcopy(oracle)
    m = ours
    m.load_state_dict(o.state_dict())
    batch = po.synthetic_batch(B, SIZE, seed=9)
    o.train(), m.train()
    o.training_step({k: v.clone() for k, v in batch.items()}, 0)
    m.training_step({k: v.cuda() for k, v in batch.items()}, 0)
    sd_o, sd_m = o.state_dict(), m.state_dict()
    for k in sd_o:
        if "running_mean" in k or "running_var" in k:
            a, b = sd_m[k].cpu().double(), sd_o[k].double()
            assert float((a - b).norm() / b.norm().clamp_min(1e-6)) < 3e-2, k
        if "num_batches_tracked" in k:
            assert int(sd_m[k]) == int(sd_o[k]), k
    o.eval(), m.eval()
    with torch.no_grad():
        lo = o.validation_step({k: v.clone() for k, v in batch.items()}, 0)["loss"].item()
        lm = m.validation_step({k: v.cuda() for k, v in batch.items()}, 0)["loss"].item()
    assert abs(lo - lm) <= 2e-2, (lo, lm)


def test_checkpoint_layout_roundtrip(warm):
    """state_dict -> {"state_dict": ...} file -> peclr_to_torchvision (reference contract, port_model.py:7-48)."""
    import torchvision

    from peclr_b200.port_model import peclr_to_torchvision

    cfg, oracle, ours = warm
    sd = ours.state_dict()
    assert list(sd.keys()) == list(oracle.state_dict().keys()) and len(sd) == 328
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "epoch=0.ckpt")
        torch.save({"state_dict": sd}, path)
        tv = torchvision.models.resnet50(weights=None)
        peclr_to_torchvision(tv, path)
        feats = [(k, v) for k, v in sd.items() if "features" in k]
        for (k_tv, v_tv), (k, v) in zip(tv.state_dict().items(), feats):
            assert torch.equal(v_tv, v.cpu()), (k_tv, k)
        # and it loads back into the oracle (reference-shaped) model
        fresh = po.OracleHybrid2Model(cfg)
        fresh.load_state_dict(torch.load(path)["state_dict"])
