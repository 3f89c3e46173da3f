"""Downstream consumer of the exported PeCLR encoder, inference on the sm_100a kernels: mirror of the reference's
``RN_25D_wMLPref`` (src/models/rn_25D_wMLPref.py:75-134) and ``ZrootMLP_ref`` (:6-72).

Same constructor (``backend_model`` in {"rn50", "rn152"}), same ``state_dict`` keys (``backend_model.conv1.weight`` ...
``backend_model.fc.bias``, ``zroot_ref.zroot_ref.{0,1,3,4,6}.*`` -- the released
``{rn50,rn152}_peclr_yt3d-fh_pt_fh_ft.pth`` checkpoints load with ``load_state_dict(checkpoint["state_dict"])``, README
"load them in the following manner"), same ``forward(img, K=None) -> {"kp3d", "zrel", "kp2d", "kp25d"}``.

The trunk runs on the tensor-core convolution kernels with eval-mode BatchNorm (running statistics), the 2048 -> 64
``fc`` on the fp32 GEMM kernel and everything after it (camera un-projection, root-depth quadratic, refinement MLP,
scale-normalised 3D keypoints) in one launch of ``peclr_rn25d_head`` (csrc/rn25d_head.cu).  Inference only: the
fine-tuning of this network is another experiment of the reference (out of scope); calling it in
training mode raises.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib, ops
from .resnet_model import ARCH, BNHolder, Block, ConvHolder, Placeholder


class _Backbone(nn.Module):
    """torchvision.models.resnet{50,152} parameter holder under torchvision's own names (conv1, bn1, layer1..4, fc)."""

    def __init__(self, arch):
        super().__init__()
        kind, counts = ARCH[arch]
        self.conv1, self.bn1 = ConvHolder(3, 64, 7, 2), BNHolder(64)
        inplanes = 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), counts)):
            blocks = []
            for j in range(n):
                blk = Block(kind, inplanes, planes, (1 if i == 0 else 2) if j == 0 else 1)
                inplanes = blk.out_channels
                blocks.append(blk)
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(inplanes, 3 * 21 + 1)  # 2D + zrel for 21 keypoints (+1: unused, kept for the checkpoints)


class _TrunkView(nn.Module):
    """The same module objects arranged the way the step engine walks a trunk (encoder.features / final_layer)."""

    def __init__(self, bb):
        super().__init__()
        enc = nn.Module()
        enc.features = nn.Sequential(bb.conv1, bb.bn1, Placeholder("ReLU"), Placeholder("MaxPool2d(3, 2, 1)"),
                                     bb.layer1, bb.layer2, bb.layer3, bb.layer4, Placeholder("AdaptiveAvgPool2d(1)"))
        enc.final_layer = nn.Sequential(bb.fc)
        self.encoder = enc


class ZrootMLP_ref(nn.Module):
    """Parameter holder of the root-depth refinement MLP (rn_25D_wMLPref.py:14-28); evaluated inside the head kernel."""

    def __init__(self):
        super().__init__()
        self.zroot_ref = nn.Sequential(
            nn.Linear(64, 128), nn.BatchNorm1d(128), nn.LeakyReLU(),
            nn.Linear(128, 128), nn.BatchNorm1d(128), nn.LeakyReLU(),
            nn.Linear(128, 1),
        )
        self.norm_bone_idx = (3, 8)
        self.register_buffer("eps", torch.tensor(1e-8), persistent=False)


class RN_25D_wMLPref(nn.Module):
    def __init__(self, backend_model="rn50"):
        super().__init__()
        if backend_model not in ("rn50", "rn152"):
            raise Exception(f"Unknown backend_model: {backend_model}")
        arch = {"rn50": "resnet50", "rn152": "resnet152"}[backend_model]
        # same default initialisation AND random-number consumption as the reference (:88-94): torchvision's
        # constructor, its new fc, then the refinement MLP
        import torchvision

        tv = getattr(torchvision.models, arch)()
        tv.fc = nn.Linear(tv.fc.in_features, 3 * 21 + 1)
        with torch.random.fork_rng(devices=[]):  # (the holders' own constructors must not advance the generator)
            self.backend_model = _Backbone(arch)
        own, src = self.backend_model.state_dict(), tv.state_dict()
        assert list(own.keys()) == list(src.keys())
        with torch.no_grad():
            for k, v in own.items():
                v.copy_(src[k])
        del tv
        self.zroot_ref = ZrootMLP_ref()
        self.register_buffer(
            "K_default",
            torch.Tensor([[388.9018310596544, 0.0, 112.0], [0.0, 388.71231836584275, 112.0], [0.0, 0.0, 1.0]]).reshape(1, 3, 3),
            persistent=False,
        )
        from .engine import StepEngine

        view = _TrunkView(self.backend_model)
        view.eval()
        self.__dict__["_view"] = view  # not a registered submodule: the state_dict keeps the reference's keys only
        self.__dict__["engine"] = StepEngine(view)

    # ---- device moves re-home the engine's flat buffers (as BaseModel does) ------------------------------------
    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, dtype=torch.float32, device=self.engine.device))
        if probe.dtype != torch.float32:
            raise TypeError("peclr_b200 keeps fp32 master weights; the compute precision is fixed by the kernels")
        self.engine.to(probe.device)
        self.zroot_ref._apply(fn)
        self._buffers["K_default"] = fn(self._buffers["K_default"])
        return self

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict)
        self.engine.weights_dirty = True
        return out

    def forward(self, img, K=None):
        if self.training:
            raise NotImplementedError("RN_25D_wMLPref runs inference only here (call .eval()); fine-tuning it is a "
                                      "different experiment of the reference")
        if K is None:
            K = self.K_default  # default camera matrix
        eng = self.engine
        self._view.eval()
        with torch.no_grad():
            feat = eng.encode(img.float())
            fc = self.backend_model.fc
            out = ops.linear_fwd(feat, fc.weight, fc.bias, ws=eng._head_ws)
            b = out.shape[0]
            K = K.to(device=out.device, dtype=torch.float32).contiguous()
            seq = self.zroot_ref.zroot_ref
            tensors = [seq[0].weight, seq[0].bias, seq[1].weight, seq[1].bias, seq[1].running_mean, seq[1].running_var,
                       seq[3].weight, seq[3].bias, seq[4].weight, seq[4].bias, seq[4].running_mean, seq[4].running_var,
                       seq[6].weight, seq[6].bias]
            ops._need_cuda(out, K, *tensors)
            ptrs = (ctypes.c_void_p * 14)(*[t.data_ptr() for t in tensors])
            kp3d = torch.empty((b, 21, 3), dtype=torch.float32, device=out.device)
            zrel = torch.empty((b, 21, 1), dtype=torch.float32, device=out.device)
            kp2d = torch.empty((b, 21, 2), dtype=torch.float32, device=out.device)
            kp25d = torch.empty((b, 21, 3), dtype=torch.float32, device=out.device)
            _lib.call("peclr_rn25d_head", out, K, K.shape[0], b, ptrs, float(seq[1].eps),
                      float(seq[2].negative_slope), kp3d, zrel, kp2d, kp25d, ops._s())
        return {"kp3d": kp3d, "zrel": zrel, "kp2d": kp2d, "kp25d": kp25d}
