"""Shared helpers of the step-level parity tests (GPU): the oracle (oracle/peclr_oracle.py, plain torch in the
reference's op order, bit-identical to the executed reference on the CPU -- tests/test_oracle_vs_reference.py) is run
in fp32 ON THE GPU with TF32 disabled as the full-size reference, where the CPU would take minutes per step.

Fixed tolerances of the bf16 tensor-core trunk against that fp32 reference, from an
oracle-warm-started state: gradient cosine >= 0.93 over all parameters, >= 0.985 for layer4 and for the projection
head; |dloss| <= 1.5e-3 at BASELINE's batch (B = 128: the loss is a mean over 256 rows; measured 2e-6 ... 3e-4),
<= 1e-2 for the small-batch configurations (B = 8 / 16: 16 / 32 rows -- the bf16 rounding noise of single rows does
not average out: over the configurations of scripts/parity_probe.py and these tests the reference's OWN arithmetic
under torch bf16 autocast measures 8e-5 ... 1.05e-2, this build 8e-5 ... 7e-3, uncorrelated with each other).  They
are constants: nothing is scaled by what the run happens to measure.  The step is bit-reproducible (ABI 3) and the
oracle's warm-up runs with deterministic cuDNN algorithms, so a configuration that passes once passes always.

How far a network is from torchvision's default initialisation decides how much of the bf16 rounding noise its
backward pass amplifies (chaotic at default init for ANY 16-bit trunk): after 100 Adam steps ResNet-152
still sits at cosine 0.87 -- for this build and for the reference under bf16 autocast alike (0.8646 vs 0.8696) --,
after 500 steps both are at 0.96.  The deep trunks are therefore warm-started for 400-500 steps (at B = 8, 64 x 64:
cheap on the GPU); profiles/parity_probe_r02.txt keeps the measured table, including the autocast reference.
"""
import contextlib

import torch

from oracle import peclr_oracle as po

TOL_DLOSS = 1.5e-3        # B = 128 (256 rows)
TOL_DLOSS_SMALL = 1e-2    # B = 8 / 16 (16 / 32 rows)
TOL_COS_ALL = 0.93
TOL_COS_TOP = 0.985  # layer4 and head
GROUPS = ("stem", "layer1", "layer2", "layer3", "layer4", "head", "all")


@contextlib.contextmanager
def strict_fp32():
    """fp32 oracle on the GPU: no TF32 in cuDNN convolutions or cuBLAS matmuls, deterministic cuDNN algorithms (the
    warm-up then yields the same weights in every run)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
           torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        yield
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
         torch.backends.cudnn.benchmark) = old


def group_of(name):
    if name.startswith("projection_head"):
        return "head"
    idx = name.split(".")[2]
    return {"0": "stem", "1": "stem", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}[idx]


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def grads_by_group(named):
    out = {}
    for n, g in named.items():
        out.setdefault(group_of(n), []).append(g.detach().double().flatten().cpu())
    out = {k: torch.cat(v) for k, v in out.items()}
    out["all"] = torch.cat([out[k] for k in GROUPS if k in out])
    return out


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


_WARM_CACHE = {}


def warm_started_oracle(cfg, steps, batch_size, size, device="cuda", lr=1e-3, seed=0):
    """torchvision-default-initialised oracle after `steps` Adam steps on structured synthetic batches (white noise
    at default init is numerically chaotic for any 16-bit trunk).  The warmed weights are cached per
    (trunk, augmentation, steps, batch, size) within the process: several tests share one warm-up."""
    key = (cfg["resnet_size"], tuple(cfg["augmentation"]), cfg["projection_head_input_dim"], steps, batch_size, size,
           lr, seed, str(device))
    torch.manual_seed(seed)
    oracle = po.OracleHybrid2Model(cfg).to(device)
    oracle.train()
    if key in _WARM_CACHE:
        oracle.load_state_dict(_WARM_CACHE[key])
        return oracle
    if steps:
        opt = torch.optim.Adam(oracle.parameters(), lr=lr)
        with strict_fp32():
            for i in range(steps):
                batch = po.synthetic_batch(batch_size, size, seed=100 + i)
                if device != "cpu":
                    batch = to_cuda(batch)
                opt.zero_grad(set_to_none=True)
                oracle.training_step(batch, i)["loss"].backward()
                opt.step()
    oracle.zero_grad(set_to_none=True)
    oracle.train_metrics, oracle.plot_params = {}, {}
    _WARM_CACHE[key] = {k: v.detach().clone() for k, v in oracle.state_dict().items()}
    return oracle


def candidate_from(oracle, cfg, cls=None):
    """The CUDA model with the oracle's weights and buffers."""
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    torch.manual_seed(0)
    ours = (cls or Hybrid2Model)(EasyDict(dict(cfg)))
    ours.load_state_dict({k: v.cpu() for k, v in oracle.state_dict().items()})
    return ours.cuda()


def oracle_step_on_gpu(oracle, batch):
    """loss / statistics / gradients of the fp32 oracle on the GPU (running statistics restored afterwards)."""
    sd = {k: v.clone() for k, v in oracle.state_dict().items()}
    oracle.train()
    oracle.zero_grad(set_to_none=True)
    oracle.train_metrics = {}
    with strict_fp32():
        out = oracle.training_step({k: v.clone() for k, v in batch.items()}, 0)
        out["loss"].backward()
    torch.cuda.synchronize()
    res = {k: float(v) for k, v in out.items()}
    grads = grads_by_group(po.named_grads(oracle))
    oracle.zero_grad(set_to_none=True)
    oracle.train_metrics, oracle.plot_params = {}, {}
    oracle.load_state_dict(sd)
    return res, grads


def autocast_envelope(oracle, batch, crop=True, rotate=True):
    """For context only (never asserted against): what the REFERENCE arithmetic itself gives when its trunk + head
    run under torch's bf16 autocast on the same weights / batch (fp32 loss chain) -- loss and per-group gradients."""
    import numpy as np

    sd = {k: v.clone() for k, v in oracle.state_dict().items()}
    oracle.train()
    oracle.zero_grad(set_to_none=True)
    x = torch.cat((batch["transformed_image1"], batch["transformed_image2"]))
    size = tuple(x.shape[-2:])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        proj = oracle.projection_head(oracle.encoder(x))
    proj = proj.float()
    cat = lambda a, b: torch.cat((batch[a], batch[b])).cpu().numpy()
    chain = po.loss_chain_numpy(proj.detach().cpu().numpy(), cat("angle_1", "angle_2") if rotate else None,
                                cat("jitter_x_1", "jitter_x_2") if crop else None,
                                cat("jitter_y_1", "jitter_y_2") if crop else None, size, crop, rotate, dtype=np.float32)
    proj.backward(torch.tensor(chain["g_p"], device=proj.device))
    torch.cuda.synchronize()
    grads = grads_by_group(po.named_grads(oracle))
    oracle.zero_grad(set_to_none=True)
    oracle.load_state_dict(sd)
    return float(chain["loss"]), grads


def candidate_step(ours, batch):
    ours.train()
    ours.zero_grad()
    out = ours.training_step(batch, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    res = {k: float(v) for k, v in out.items()}
    grads = grads_by_group({n: p.grad for n, p in ours.named_parameters() if not n.startswith("encoder.final_layer")})
    return res, grads


def report_and_check(tag, got, got_g, ref, ref_g, check_stats=True, envelope=None, tol=None):
    """tol: (dloss, cos_all, cos_top) constants of the calling test; default = the SURVEY 8(d)(iii) triple."""
    tol_dloss, tol_all, tol_top = tol or (TOL_DLOSS, TOL_COS_ALL, TOL_COS_TOP)
    dloss = abs(got["loss"] - ref["loss"])
    cosines = {k: cos(got_g[k], ref_g[k]) for k in ref_g}
    env_loss, env_g = envelope if envelope is not None else (None, None)
    print("\n[parity %s] loss ours %.6f oracle(fp32) %.6f |d| %.2e%s  (tolerances: |d| %.1e, cos all %.3f, layer4/head "
          "%.3f)" % (tag, got["loss"], ref["loss"], dloss,
                    "" if env_loss is None else " (reference under bf16 autocast |d| %.2e)" % abs(env_loss - ref["loss"]),
                    tol_dloss, tol_all, tol_top))
    for k in GROUPS:
        if k in ref_g:
            rel = float((got_g[k] - ref_g[k]).norm() / ref_g[k].norm().clamp_min(1e-300))
            env = "" if env_g is None else "  [reference under bf16 autocast: cos %.4f]" % cos(env_g[k], ref_g[k])
            print("  %-7s |g| %.3e  cos %.4f  rel-L2 %.3f%s" % (k, float(ref_g[k].norm()), cosines[k], rel, env))
    assert set(got) == set(ref)
    assert dloss <= tol_dloss, dloss
    if check_stats:
        for k in ref:
            if k != "loss":
                assert abs(got[k] - ref[k]) <= 3e-2 * (abs(ref[k]) + 0.05), (k, got[k], ref[k])
    assert cosines["all"] >= tol_all, cosines
    assert cosines["layer4"] >= tol_top and cosines["head"] >= tol_top, cosines
    return dloss, cosines
