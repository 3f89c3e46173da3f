"""Shared helpers of the step-level parity tests (GPU): the oracle (oracle/peclr_oracle.py, plain torch in the
reference's op order, bit-identical to the executed reference on the CPU -- tests/test_oracle_vs_reference.py) is run
in fp32 ON THE GPU with TF32 disabled as the full-size reference, where the CPU would take minutes per step.

Fixed tolerances of the bf16 tensor-core trunk against that fp32 reference, from an
oracle-warm-started state: |dloss| <= 1.5e-3, gradient cosine >= 0.93 over all parameters, >= 0.985 for layer4 and
for the projection head.  They are constants: nothing is scaled by what the run happens to measure.
"""
import contextlib

import torch

from oracle import peclr_oracle as po

TOL_DLOSS = 1.5e-3
TOL_COS_ALL = 0.93
TOL_COS_TOP = 0.985  # layer4 and head
GROUPS = ("stem", "layer1", "layer2", "layer3", "layer4", "head", "all")


@contextlib.contextmanager
def strict_fp32():
    """fp32 oracle on the GPU: no TF32 in cuDNN convolutions or cuBLAS matmuls."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


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


def warm_started_oracle(cfg, steps, batch_size, size, device="cuda", lr=1e-3, seed=0):
    """torchvision-default-initialised oracle after `steps` Adam steps on structured synthetic batches (white noise
    at default init is numerically chaotic for any 16-bit trunk)."""
    torch.manual_seed(seed)
    oracle = po.OracleHybrid2Model(cfg).to(device)
    oracle.train()
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


def candidate_step(ours, batch):
    ours.train()
    ours.zero_grad()
    out = ours.training_step(batch, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    res = {k: float(v) for k, v in out.items()}
    grads = grads_by_group({n: p.grad for n, p in ours.named_parameters() if not n.startswith("encoder.final_layer")})
    return res, grads


def report_and_check(tag, got, got_g, ref, ref_g, check_stats=True):
    dloss = abs(got["loss"] - ref["loss"])
    cosines = {k: cos(got_g[k], ref_g[k]) for k in ref_g}
    print("\n[parity %s] loss ours %.6f oracle(fp32) %.6f |d| %.2e" % (tag, got["loss"], ref["loss"], dloss))
    for k in GROUPS:
        if k in ref_g:
            rel = float((got_g[k] - ref_g[k]).norm() / ref_g[k].norm().clamp_min(1e-300))
            print("  %-7s |g| %.3e  cos %.4f  rel-L2 %.3f" % (k, float(ref_g[k].norm()), cosines[k], rel))
    assert set(got) == set(ref)
    assert dloss <= TOL_DLOSS, dloss
    if check_stats:
        for k in ref:
            if k != "loss":
                assert abs(got[k] - ref[k]) <= 3e-2 * (abs(ref[k]) + 0.05), (k, got[k], ref[k])
    assert cosines["all"] >= TOL_COS_ALL, cosines
    assert cosines["layer4"] >= TOL_COS_TOP and cosines["head"] >= TOL_COS_TOP, cosines
    return dloss, cosines
