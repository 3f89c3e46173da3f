"""Downstream consumer of the exported encoder on the B200: RN_25D_wMLPref inference
(reference src/models/rn_25D_wMLPref.py:75-134) on the CUDA trunk + the fused head kernel, against the oracle
restatement (oracle/rn25d_oracle.py, bit-identical to the reference module -- tests/test_oracle_vs_reference.py) run in
fp32 on the GPU.

Tolerances: the head kernel alone (fp32, same backbone output on both sides) 1e-5 relative; the whole network with the
bf16 tensor-core trunk rel-L2 <= 1.5e-2 on every output (eval-mode BatchNorm; measured 0.2e-2 ... 0.7e-2).
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _oracle(backend, seed=0, settle=6):
    """Default-initialised oracle whose BatchNorm running statistics have seen a few batches (eval mode with the
    initial (0, 1) statistics would be a degenerate test)."""
    import parity_util as pu
    from oracle import peclr_oracle as po
    from oracle.rn25d_oracle import OracleRN25D

    torch.manual_seed(seed)
    oracle = OracleRN25D(backend).cuda()
    with torch.no_grad():
        # default-initialised fc rows sum 2048 post-ReLU features with random signs: outputs near zero by cancellation,
        # so their RELATIVE error says nothing.  A positive component makes them keypoint-like magnitudes (tens of
        # pixels), as trained checkpoints have.
        oracle.backend_model.fc.weight.add_(0.03)
    oracle.train()
    with torch.no_grad(), pu.strict_fp32():
        for i in range(settle):
            x = po.synthetic_batch(8, 224, seed=700 + i)["transformed_image1"].cuda()
            for m in oracle.modules():
                if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                    m.momentum = 0.5
            oracle(x)
    oracle.eval()
    return oracle


def test_head_kernel_vs_oracle():
    import parity_util as pu
    from peclr_b200 import _lib, ops

    oracle = _oracle("rn50", settle=2)
    g = torch.Generator(device="cuda").manual_seed(2)
    for b, per_sample_k in ((1, False), (5, True), (64, False)):
        out = torch.randn(b, 64, device="cuda", generator=g)
        out[:, 0:63:3] = out[:, 0:63:3] * 40 + 112  # pixel-like 2D coordinates
        out[:, 1:63:3] = out[:, 1:63:3] * 40 + 112
        K = None
        if per_sample_k:
            K = oracle.K_default.repeat(b, 1, 1).clone()
            K[:, 0, 0] += torch.arange(b, device="cuda") * 3.0
            K[:, 1, 2] -= torch.arange(b, device="cuda") * 1.5
        with torch.no_grad(), pu.strict_fp32():
            want = oracle.head(out, K)
        seq = oracle.zroot_ref.zroot_ref
        tensors = [seq[0].weight, seq[0].bias, seq[1].weight, seq[1].bias, seq[1].running_mean, seq[1].running_var,
                   seq[3].weight, seq[3].bias, seq[4].weight, seq[4].bias, seq[4].running_mean, seq[4].running_var,
                   seq[6].weight, seq[6].bias]
        ptrs = (ctypes.c_void_p * 14)(*[t.data_ptr() for t in tensors])
        kk = (K if K is not None else oracle.K_default).contiguous()
        got = {"kp3d": torch.empty(b, 21, 3, device="cuda"), "zrel": torch.empty(b, 21, 1, device="cuda"),
               "kp2d": torch.empty(b, 21, 2, device="cuda"), "kp25d": torch.empty(b, 21, 3, device="cuda")}
        _lib.call("peclr_rn25d_head", out, kk, kk.shape[0], b, ptrs, 1e-5, 0.01, got["kp3d"], got["zrel"], got["kp2d"],
                  got["kp25d"], ops._s())
        torch.cuda.synchronize()
        for k in want:
            assert got[k].shape == want[k].shape, k
            assert rel_l2(got[k], want[k]) < 1e-5, (b, k, rel_l2(got[k], want[k]))
        assert float(got["zrel"][:, 0].abs().max()) == 0 and float(got["kp25d"][:, 0, 2].abs().max()) == 0


@pytest.mark.parametrize("backend", ["rn50", "rn152"])
def test_rn25d_inference_vs_fp32_oracle(backend):
    import parity_util as pu
    from oracle import peclr_oracle as po
    from peclr_b200.rn_25D_wMLPref import RN_25D_wMLPref

    oracle = _oracle(backend)
    ours = RN_25D_wMLPref(backend_model=backend)
    ours.load_state_dict({k: v.cpu() for k, v in oracle.state_dict().items()})  # README: load_state_dict(ckpt["state_dict"])
    ours.eval()
    ours.to(torch.device("cuda"))
    K = oracle.K_default.repeat(3, 1, 1).clone()
    K[1, 0, 0] = 371.0
    for n, kk in ((1, None), (3, K), (8, None)):
        x = po.synthetic_batch(max(n, 2), 224, seed=800 + n)["transformed_image2"][:n].cuda()
        with torch.no_grad(), pu.strict_fp32():
            want = oracle(x, kk)
        got = ours(x, kk)
        torch.cuda.synchronize()
        assert set(got) == {"kp3d", "zrel", "kp2d", "kp25d"}
        errs = {k: rel_l2(got[k], want[k]) for k in want}
        print("\n[rn25d %s n=%d] rel-L2 vs fp32 oracle" % (backend, n), {k: round(v, 4) for k, v in errs.items()})
        for k in want:
            assert got[k].shape == want[k].shape and errs[k] < 1.5e-2, (k, errs[k])
    ours.train()
    with pytest.raises(NotImplementedError):
        ours(x)
    # state_dict keeps the reference's keys (what a released checkpoint holds)
    assert list(ours.state_dict().keys()) == list(oracle.state_dict().keys())
