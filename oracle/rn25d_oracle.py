"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement (oracle) of the downstream 2.5D hand-pose network that consumes
the exported PeCLR encoders: ``RN_25D_wMLPref`` and ``ZrootMLP_ref`` of the reference
(src/models/rn_25D_wMLPref.py:6-72 and :75-134), same module / state_dict layout, same op order.

Nothing in the product path may import this module; tests use it as the checker.  Pinning: in the build container
the restatement is compared with the reference module itself, imported from /root/reference (it only needs torch and
torchvision) -- identical state_dict keys / initial values for the same seed and bit-identical outputs
(tests/test_oracle_vs_reference.py::test_rn25d_oracle_equals_reference).  The reference ships no tests or golden
vectors for it.
"""
import torch
import torch.nn as nn
from torchvision import models


class OracleZrootMLP(nn.Module):
    """rn_25D_wMLPref.py:6-72: zroot_ref = zroot_est + mlp(2D, zrel, zroot_est)."""

    def __init__(self):
        super().__init__()
        self.zroot_ref = nn.Sequential(
            nn.Linear(64, 128), nn.BatchNorm1d(128), nn.LeakyReLU(),
            nn.Linear(128, 128), nn.BatchNorm1d(128), nn.LeakyReLU(),
            nn.Linear(128, 1),
        )
        self.norm_bone_idx = (3, 8)
        self.register_buffer("eps", torch.tensor(1e-8), persistent=False)

    def forward(self, kp3d_unnorm, zrel, K):
        eps = self.eps
        m, n = self.norm_bone_idx
        X_m, Y_m = kp3d_unnorm[:, m:m + 1, 0:1], kp3d_unnorm[:, m:m + 1, 1:2]
        X_n, Y_n = kp3d_unnorm[:, n:n + 1, 0:1], kp3d_unnorm[:, n:n + 1, 1:2]
        zrel_m, zrel_n = zrel[:, m:m + 1], zrel[:, n:n + 1]
        # scale-normalised root depth from the reference bone (Iqbal et al. 2018, eq. 6-7) -- :38-58
        a = (X_n - X_m) ** 2 + (Y_n - Y_m) ** 2
        b = 2 * (zrel_n * (X_n ** 2 + Y_n ** 2 - X_n * X_m - Y_n * Y_m)
                 + zrel_m * (X_m ** 2 + Y_m ** 2 - X_n * X_m - Y_n * Y_m))
        c = ((X_n * zrel_n - X_m * zrel_m) ** 2 + (Y_n * zrel_n - Y_m * zrel_m) ** 2 + (zrel_n - zrel_m) ** 2 - 1)
        d = (b ** 2) - (4 * a * c)
        a = torch.max(eps, a)
        d = torch.max(eps, d)
        zroot = ((-b + torch.sqrt(d)) / (2 * a)).detach()
        zroot = torch.clamp(zroot, 4.0, 50.0)  # :60
        mlp_input = torch.cat((zrel.reshape(-1, 21), kp3d_unnorm[..., :2].reshape(-1, 42), zroot.reshape(-1, 1)), dim=1)
        return zroot + self.zroot_ref(mlp_input).reshape(zroot.shape)


class OracleRN25D(nn.Module):
    """rn_25D_wMLPref.py:75-134."""

    def __init__(self, backend_model="rn50"):
        super().__init__()
        if backend_model == "rn50":
            model_func = models.resnet50
        elif backend_model == "rn152":
            model_func = models.resnet152
        else:
            raise Exception(f"Unknown backend_model: {backend_model}")
        backend = model_func()
        backend.fc = nn.Linear(backend.fc.in_features, 3 * 21 + 1)
        self.backend_model = backend
        self.zroot_ref = OracleZrootMLP()
        self.register_buffer(
            "K_default",
            torch.Tensor([[388.9018310596544, 0.0, 112.0], [0.0, 388.71231836584275, 112.0], [0.0, 0.0, 1.0]]).reshape(1, 3, 3),
            persistent=False,
        )

    def head(self, out, K=None):
        """Everything after the backbone (:109-134) on a given backbone output [B, 64]."""
        if K is None:
            K = self.K_default
        out = out.clone()
        kp25d = out[:, :-1].view(-1, 21, 3)
        kp2d = kp25d[..., :2]
        zrel = kp25d[..., 2:3]
        zrel[:, 0] = 0  # zrel of the root is 0 (written through the views into kp25d)
        kp2d_h = torch.cat((kp2d, torch.ones((kp2d.shape[0], 21, 1), device=K.device)), dim=2)
        kp3d_unnorm = torch.matmul(kp2d_h, K.inverse().transpose(1, 2))
        zroot = self.zroot_ref(kp3d_unnorm, zrel, K)
        kp3d = kp3d_unnorm * (zrel + zroot)
        return {"kp3d": kp3d, "zrel": zrel, "kp2d": kp2d, "kp25d": kp25d}

    def forward(self, img, K=None):
        return self.head(self.backend_model(img), K)
