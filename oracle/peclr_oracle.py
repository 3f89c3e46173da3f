"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the PeCLR pre-training step.

Nothing in the product path (``peclr_b200/``, ``src/``) may import this module.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU baseline.

What it restates (reference = dahiyaaneesh/peclr, mounted at /root/reference in the
build container only):

* the loss chain  normalise -> inverse translate -> inverse rotate -> normalise ->
  NT-Xent   (src/models/unsupervised/hybrid2_model.py:27-90,
  src/models/utils.py:154-186, 271-346) twice: once with the same torch ops in the same
  order (``*_torch`` functions, used to build the full-step oracle model), once as a
  closed-form numpy forward + hand-derived backward (``loss_chain_numpy``) in the
  requested float width;
* the model shell (src/models/resnet_model.py:6-56, simclr_model.py:10-76,
  base_model.py:13-127) as ``OracleHybrid2Model`` on top of torchvision's ResNet;
* the optimiser arithmetic that lives in the un-vendored dependency
  pytorch-lightning-bolts==0.2.2 (requirements.txt:105): ``LARSWrapper`` and
  ``LinearWarmupCosineAnnealingLR`` -- restated from the library's published
  algorithm; call sites base_model.py:90-98.

Pinning status: the loss chain, model shell, step and checkpoint layout are pinned
against the *executed* reference in the build container (tests/test_oracle_vs_reference.py,
fixtures in tests/golden/ made by oracle/make_golden.py).  The reference ships no tests
or golden vectors of its own, and pl_bolts is not available
offline, so ``LARSWrapper`` / ``LinearWarmupCosineAnnealingLR`` are **parity unpinned**.
"""
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# 1. closed-form numpy loss chain (forward + backward)
# --------------------------------------------------------------------------------------


def loss_chain_numpy(
    p: np.ndarray,
    angle: Optional[np.ndarray],
    jitter_x: Optional[np.ndarray],
    jitter_y: Optional[np.ndarray],
    image_hw: Tuple[int, int],
    crop: bool,
    rotate: bool,
    temperature: float = 0.5,
    dtype=np.float64,
    want_grad: bool = True,
) -> Dict[str, np.ndarray]:
    """Loss chain of Hybrid2Model.contrastive_step on the projection-head output.

    p         (2B, D) projection-head output, rows [view1 ; view2], D even, read as D/2
              interleaved (x, y) points                      hybrid2_model.py:38
    angle     (2B,) degrees, float64 (angle_1 ; angle_2)     hybrid2_model.py:78
    jitter_*  (2B,) integer pixels (jitter_*_1 ; jitter_*_2) hybrid2_model.py:59-72
    image_hw  image.size()[-2:]; x is divided by [0], y by [1]  hybrid2_model.py:34-35,61-70

    Returns loss, z (2B, D) and, if want_grad, g_p = dloss/dp.
    """
    p = np.asarray(p, dtype=dtype)
    n, d = p.shape
    b = n // 2
    eps = dtype(1e-12)
    # F.normalize (hybrid2_model.py:48-49): x / max(||x||, eps)
    pn = np.sqrt((p * p).sum(1, dtype=dtype))
    pn_c = np.maximum(pn, eps)
    u = p / pn_c[:, None]
    pts = u.reshape(n, d // 2, 2).copy()
    if crop:
        # hybrid2_model.py:59-74 + utils.py:325-346 ; max/min are detached.
        # int64 / python float -> default dtype (f32 in the fp32 run, f64 in the fp64 run)
        tx = -(np.asarray(jitter_x).astype(dtype) / dtype(image_hw[0]))
        ty = -(np.asarray(jitter_y).astype(dtype) / dtype(image_hw[1]))
        rng_x = pts[:, :, 0].max(1) - pts[:, :, 0].min(1)
        rng_y = pts[:, :, 1].max(1) - pts[:, :, 1].min(1)
        pts[:, :, 0] += (tx * rng_x)[:, None]
        pts[:, :, 1] += (ty * rng_y)[:, None]
    if rotate:
        # hybrid2_model.py:76-80 + utils.py:271-321 ; trig in float64, matrix stored in
        # the default dtype; centre = detached mean of the (translated) points.
        a = -np.asarray(angle, dtype=np.float64) * np.pi / 180.0
        cx = pts[:, :, 0].mean(1, dtype=dtype).astype(np.float64)
        cy = pts[:, :, 1].mean(1, dtype=dtype).astype(np.float64)
        al64, be64 = np.cos(a), np.sin(a)
        al, be = al64.astype(dtype), be64.astype(dtype)
        offx = ((1 - al64) * cx - be64 * cy).astype(dtype)
        offy = ((1 - al64) * cy + be64 * cx).astype(dtype)
        x, y = pts[:, :, 0].copy(), pts[:, :, 1].copy()
        pts[:, :, 0] = al[:, None] * x + be[:, None] * y + offx[:, None]
        pts[:, :, 1] = -be[:, None] * x + al[:, None] * y + offy[:, None]
    else:
        al = np.ones(n, dtype=dtype)
        be = np.zeros(n, dtype=dtype)
    r = pts.reshape(n, d)
    rn = np.sqrt((r * r).sum(1, dtype=dtype))
    rn_c = np.maximum(rn, eps)
    z = r / rn_c[:, None]
    # NT-Xent (utils.py:154-186): self excluded, positive included, no max-subtraction.
    t = dtype(temperature)
    s = (z @ z.T) / t
    e = np.exp(s)
    np.fill_diagonal(e, 0)
    neg = e.sum(1, dtype=dtype)
    pos_idx = (np.arange(n) + b) % n
    spos = (z * z[pos_idx]).sum(1, dtype=dtype) / t
    loss = -(spos - np.log(neg)).mean(dtype=dtype)
    out = {"loss": dtype(loss), "z": z, "neg": neg}
    if not want_grad:
        return out
    pm = e / neg[:, None]
    g = pm.copy()
    g[np.arange(n), pos_idx] -= 1
    g /= t * n
    gz = (g + g.T) @ z
    gr = (gz - z * (z * gz).sum(1, keepdims=True)) / rn_c[:, None]
    gr = np.where((rn > eps)[:, None], gr, gz / eps)
    gp3 = gr.reshape(n, d // 2, 2)
    gx = al[:, None] * gp3[:, :, 0] - be[:, None] * gp3[:, :, 1]
    gy = be[:, None] * gp3[:, :, 0] + al[:, None] * gp3[:, :, 1]
    gu = np.stack([gx, gy], axis=2).reshape(n, d)
    gp = (gu - u * (u * gu).sum(1, keepdims=True)) / pn_c[:, None]
    gp = np.where((pn > eps)[:, None], gp, gu / eps)
    out["g_p"] = gp.astype(dtype)
    return out


def projection_stats_numpy(p: np.ndarray, name: str) -> Dict[str, float]:
    """hybrid2_model.py:92-106 on one view's (B, D/2, 2) projections; lower median."""
    b, m, _ = p.shape
    srt = np.sort(p, axis=1)
    med = srt[:, (m - 1) // 2, :]
    out = {}
    for c, cname in enumerate("xy"):
        out[f"{name}{cname}_mean"] = float(p[:, :, c].mean(1).mean())
        out[f"{name}{cname}_median"] = float(med[:, c].mean())
        out[f"{name}{cname}_min"] = float(p[:, :, c].min(1).mean())
        out[f"{name}{cname}_max"] = float(p[:, :, c].max(1).mean())
    return out


# --------------------------------------------------------------------------------------
# 2. learning-rate schedule and LARS-Adam arithmetic (pl_bolts 0.2.2, unpinned)
# --------------------------------------------------------------------------------------


def warmup_cosine_lr(step: int, base_lr: float, warmup: int, max_steps: int,
                     start_lr: float = 0.0, eta_min: float = 0.0) -> float:
    """Closed form of pl_bolts LinearWarmupCosineAnnealingLR (value after `step` calls to
    scheduler.step(); step 0 = value at construction)."""
    if step < warmup:
        return start_lr + step * (base_lr - start_lr) / (warmup - 1)
    return eta_min + 0.5 * (base_lr - eta_min) * (
        1 + math.cos(math.pi * (step - warmup) / (max_steps - warmup))
    )


def lars_adam_step_numpy(p, g, m, v, step, lr, wd, lars=True, eta=0.02, clip=True,
                         lars_eps=1e-8, beta1=0.9, beta2=0.999, adam_eps=1e-8):
    """One LARSWrapper(Adam).step() on a single tensor, float32 arithmetic as torch does.

    step is the 1-based Adam step count after this update.  With ``lars`` the group's
    weight decay is folded into the gradient by the wrapper and Adam sees wd = 0
    (base_model.py:90-91, pl_bolts lars_scheduling.py); without it Adam applies L2 wd.
    Returns (p, m, v) updated copies.
    """
    f = np.float32
    p, g, m, v = (np.array(x, dtype=f) for x in (p, g, m, v))
    if lars:
        pn = f(np.sqrt((p.astype(np.float64) ** 2).sum()))
        gn = f(np.sqrt((g.astype(np.float64) ** 2).sum()))
        if pn != 0 and gn != 0:
            new_lr = f(f(eta) * pn) / f(f(gn + f(pn * f(wd))) + f(lars_eps))
            if clip:
                new_lr = min(f(new_lr / f(lr)) if lr != 0 else f(np.inf), f(1.0))
            g = f(new_lr) * (g + f(wd) * p)
    else:
        g = g + f(wd) * p
    m = f(beta1) * m + f(1 - beta1) * g
    v = f(beta2) * v + f(1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(adam_eps)
    p = p - f(step_size) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


# --------------------------------------------------------------------------------------
# 3. torch restatement (same ops, same order as the reference) -- full-step oracle
# --------------------------------------------------------------------------------------

import torch  # noqa: E402
from torch import Tensor, nn  # noqa: E402
from torch.nn import functional as F  # noqa: E402


def vanila_contrastive_loss(z1: Tensor, z2: Tensor, temperature: float = 0.5) -> Tensor:
    """NT-Xent as in src/models/utils.py:154-186."""
    z = torch.cat([z1, z2], dim=0)
    n = len(z)
    sim = torch.exp(torch.mm(z, z.t().contiguous()) / temperature)
    off_diag = ~torch.eye(n, device=sim.device).bool()
    neg = sim.masked_select(off_diag).view(n, -1).sum(dim=-1)
    pos = torch.exp(torch.sum(z1 * z2, dim=-1) / temperature)
    pos = torch.cat([pos, pos], dim=0)
    return -torch.log(pos / neg).mean()


def get_rotation_2D_matrix(angle: Tensor, center_x: Tensor, center_y: Tensor, scale) -> Tensor:
    """src/models/utils.py:271-298 -- (n,3,2) matrix allocated on CPU in the default dtype."""
    angle = angle * np.pi / 180
    alpha = scale * torch.cos(angle)
    beta = scale * torch.sin(angle)
    rot = torch.zeros((len(angle), 3, 2))
    rot[:, :, 0] = torch.stack([alpha, beta, (1 - alpha) * center_x - beta * center_y], dim=1)
    rot[:, :, 1] = torch.stack([-beta, alpha, (1 - alpha) * center_y + beta * center_x], dim=1)
    return rot


def rotate_encoding(encoding: Tensor, angle: Tensor) -> Tensor:
    """src/models/utils.py:301-321 (in place on `encoding`)."""
    centre = torch.mean(encoding.detach(), 1)
    rot = get_rotation_2D_matrix(angle, centre[:, 0], centre[:, 1], scale=1.0).to(encoding.device)
    homog = torch.cat((encoding[..., :2], torch.ones_like(encoding[..., -1:])), dim=2)
    encoding[..., :2] = torch.bmm(homog, rot)
    return encoding


def translate_encodings(encoding: Tensor, translate_x: Tensor, translate_y: Tensor) -> Tensor:
    """src/models/utils.py:325-346 (in place on `encoding`)."""
    hi = torch.max(encoding.detach(), dim=1).values
    lo = torch.min(encoding.detach(), dim=1).values
    encoding[..., 0] += (translate_x * (hi[:, 0] - lo[:, 0])).view((-1, 1))
    encoding[..., 1] += (translate_y * (hi[:, 1] - lo[:, 1])).view((-1, 1))
    return encoding


class AttrDict(dict):
    """Stand-in for easydict.EasyDict (not installed)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class LARSWrapper:
    """pl_bolts 0.2.2 optimizers/lars_scheduling.py restated (unpinned)."""

    def __init__(self, optimizer, eta=0.02, clip=True, eps=1e-8):
        self.optim = optimizer
        self.eta, self.clip, self.eps = eta, clip, eps

    param_groups = property(lambda self: self.optim.param_groups)
    state = property(lambda self: self.optim.state)
    defaults = property(lambda self: self.optim.defaults)

    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, sd):
        return self.optim.load_state_dict(sd)

    def zero_grad(self, *a, **k):
        return self.optim.zero_grad(*a, **k)

    @torch.no_grad()
    def step(self, closure=None):
        saved = []
        for group in self.optim.param_groups:
            wd = group.get("weight_decay", 0)
            saved.append(wd)
            group["weight_decay"] = 0
            for p in group["params"]:
                if p.grad is not None:
                    self.update_p(p, group, wd)
        self.optim.step(closure=closure)
        for group, wd in zip(self.optim.param_groups, saved):
            group["weight_decay"] = wd

    def update_p(self, p, group, weight_decay):
        p_norm = torch.norm(p.data)
        g_norm = torch.norm(p.grad.data)
        if p_norm != 0 and g_norm != 0:
            new_lr = (self.eta * p_norm) / (g_norm + p_norm * weight_decay + self.eps)
            if self.clip:
                lr = group["lr"]
                ratio = new_lr / lr if lr != 0 else torch.full_like(new_lr, float("inf"))
                new_lr = torch.clamp(ratio, max=1.0)
            p.grad.data += weight_decay * p.data
            p.grad.data *= new_lr


class LinearWarmupCosineAnnealingLR:
    """pl_bolts 0.2.2 optimizers/lr_scheduler.py restated (recursive form, unpinned)."""

    def __init__(self, optimizer, warmup_epochs, max_epochs, warmup_start_lr=0.0, eta_min=0.0):
        self.optimizer = optimizer
        self.warmup_epochs, self.max_epochs = warmup_epochs, max_epochs
        self.warmup_start_lr, self.eta_min = warmup_start_lr, eta_min
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.last_epoch = -1
        self.step()

    def get_lr(self) -> List[float]:
        e, w, m = self.last_epoch, self.warmup_epochs, self.max_epochs
        groups = self.optimizer.param_groups
        if e == 0:
            return [self.warmup_start_lr] * len(self.base_lrs)
        if e < w:
            return [g["lr"] + (b - self.warmup_start_lr) / (w - 1) for b, g in zip(self.base_lrs, groups)]
        if e == w:
            return list(self.base_lrs)
        if (e - 1 - m) % (2 * (m - w)) == 0:
            return [g["lr"] + (b - self.eta_min) * (1 - math.cos(math.pi / (m - w))) / 2
                    for b, g in zip(self.base_lrs, groups)]
        return [
            (1 + math.cos(math.pi * (e - w) / (m - w)))
            / (1 + math.cos(math.pi * (e - w - 1) / (m - w)))
            * (g["lr"] - self.eta_min) + self.eta_min
            for g in groups
        ]

    def step(self):
        self.last_epoch += 1
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]


class OracleResNetModel(nn.Module):
    """src/models/resnet_model.py:6-56 in mode="pretraining" (pretrained forced False)."""

    def __init__(self, resnet_size: str):
        super().__init__()
        import torchvision.models as models

        model = getattr(models, "resnet" + str(resnet_size))(weights=None, norm_layer=nn.BatchNorm2d)
        self.features = nn.Sequential(
            model.conv1, model.bn1, model.relu, model.maxpool,
            model.layer1, model.layer2, model.layer3, model.layer4,
            nn.AdaptiveAvgPool2d(output_size=(1, 1)),
        )
        self.final_layer = nn.Sequential(nn.Linear(model.fc.in_features, 21 * 3 + 1))

    def forward(self, x):
        return self.features(x).flatten(start_dim=1)


class _TrainerStub:
    def __init__(self, world_size=1, max_epochs=100):
        self.world_size, self.max_epochs = world_size, max_epochs


class OracleHybrid2Model(nn.Module):
    """BaseModel + SimCLR + Hybrid2Model (base_model.py:13-127, simclr_model.py:10-76,
    hybrid2_model.py:16-106) as one plain nn.Module."""

    def __init__(self, config):
        super().__init__()
        self.encoder = OracleResNetModel(config["resnet_size"])
        self.config = config
        self.train_metrics: Dict[str, Tensor] = {}
        self.plot_params = {}
        self.trainer = _TrainerStub()
        self.projection_head = nn.Sequential(
            nn.Linear(config["projection_head_input_dim"], config["projection_head_hidden_dim"], bias=True),
            nn.BatchNorm1d(config["projection_head_hidden_dim"]),
            nn.ReLU(),
            nn.Linear(config["projection_head_hidden_dim"], config["output_dim"], bias=False),
        )

    # -- hybrid2_model.py:92-106
    @staticmethod
    def get_projection_stats(projection: Tensor, name: str) -> Dict[str, Tensor]:
        mean = torch.mean(projection, dim=1)
        med = torch.median(projection, dim=1).values
        lo = torch.min(projection, dim=1).values
        hi = torch.max(projection, dim=1).values
        out = {}
        for c, cname in enumerate("xy"):
            out[f"{name}{cname}_mean"] = torch.mean(mean, dim=0)[c]
            out[f"{name}{cname}_median"] = torch.mean(med, dim=0)[c]
            out[f"{name}{cname}_min"] = torch.mean(lo, dim=0)[c]
            out[f"{name}{cname}_max"] = torch.mean(hi, dim=0)[c]
        return out

    # -- hybrid2_model.py:27-85
    def get_transformed_projections(self, batch: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
        x = torch.cat((batch["transformed_image1"], batch["transformed_image2"]), dim=0)
        shape1 = batch["transformed_image1"].size()[-2:]
        shape2 = batch["transformed_image2"].size()[-2:]
        b = int(len(x) / 2)
        proj = self.projection_head(self.encoder(x)).view((b * 2, -1, 2))
        stat1 = self.get_projection_stats(proj[:b].detach(), "proj1")
        stat2 = self.get_projection_stats(proj[b:].detach(), "proj2")
        proj = proj.view((b * 2, -1))
        proj = torch.cat([F.normalize(proj[:b]), F.normalize(proj[b:])], dim=0).view((b * 2, -1, 2))
        self.train_metrics = {**self.train_metrics, **stat1, **stat2}
        if "crop" in self.config["augmentation"]:
            jx = torch.cat((batch["jitter_x_1"] / float(shape1[0]), batch["jitter_x_2"] / float(shape2[0])), dim=0)
            jy = torch.cat((batch["jitter_y_1"] / float(shape1[1]), batch["jitter_y_2"] / float(shape2[1])), dim=0)
            proj = translate_encodings(proj, -jx, -jy)
        if "rotate" in self.config["augmentation"]:
            angles = torch.cat((batch["angle_1"], batch["angle_2"]), dim=0)
            proj = rotate_encoding(proj, -angles)
        proj = proj.view((b * 2, -1))
        return F.normalize(proj[:b]), F.normalize(proj[b:])

    def contrastive_step(self, batch):
        z1, z2 = self.get_transformed_projections(batch)
        return vanila_contrastive_loss(z1, z2)

    # -- simclr_model.py:59-76
    def training_step(self, batch, batch_idx):
        loss = self.contrastive_step(batch)
        self.train_metrics = {**self.train_metrics, "loss": loss}
        self.plot_params = {
            "image1": batch["transformed_image1"],
            "image2": batch["transformed_image2"],
            "params": {k: v for k, v in batch.items() if "image" not in k},
        }
        return self.train_metrics

    def validation_step(self, batch, batch_idx):
        return {"loss": self.contrastive_step(batch)}

    # -- base_model.py:30-104
    def exclude_from_wt_decay(self, named_params, weight_decay, skip_list=("bias", "bn")):
        decayed, excluded = [], []
        for name, prm in named_params:
            if not prm.requires_grad:
                continue
            (excluded if any(s in name for s in skip_list) else decayed).append(prm)
        return [{"params": decayed, "weight_decay": weight_decay},
                {"params": excluded, "weight_decay": 0.0}]

    def setup(self, stage: str = "fit"):
        self.train_iters_per_epoch = self.config["num_samples"] // (
            self.trainer.world_size * self.config["batch_size"])

    def configure_optimizers(self):
        cfg = self.config
        groups = self.exclude_from_wt_decay(self.named_parameters(), cfg["opt_weight_decay"])
        opt = torch.optim.Adam(groups, lr=cfg["lr"] * math.sqrt(cfg["batch_size"] * cfg["num_of_mini_batch"]))
        warm = cfg["warmup_epochs"] * self.train_iters_per_epoch // cfg["num_of_mini_batch"]
        epochs = cfg.get("lr_max_epochs") or self.trainer.max_epochs
        max_steps = epochs * self.train_iters_per_epoch // cfg["num_of_mini_batch"]
        if cfg["optimizer"] == "LARS":
            opt = LARSWrapper(opt)
            sched = LinearWarmupCosineAnnealingLR(opt, warmup_epochs=warm, max_epochs=max_steps,
                                                  warmup_start_lr=0, eta_min=0)
        else:
            sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=max_steps)
        return [opt], [{"scheduler": sched, "interval": "step", "frequency": 1}]


def default_config(resnet_size="50", batch_size=128, num_samples=32768, augmentation=("crop", "rotate"),
                   optimizer="LARS", num_of_mini_batch=1, **over) -> AttrDict:
    """src/experiments/config/hybrid2_config.json + update_model_params (experiments/utils.py:608-615)."""
    cfg = AttrDict(
        batch_size=batch_size, lr=1e-4, opt_weight_decay=1e-6, output_dim=128,
        projection_head_hidden_dim=512, projection_head_input_dim=2048, warmup_epochs=10,
        num_of_mini_batch=num_of_mini_batch, augmentation=list(augmentation), optimizer=optimizer,
        resnet_size=str(resnet_size), num_samples=num_samples,
    )
    cfg.update(over)
    return cfg


# --------------------------------------------------------------------------------------
# 4. synthetic two-view batches in the reference's batch-dict schema
# --------------------------------------------------------------------------------------


def synthetic_batch(batch_size: int, size: int, seed: int = 5, structured: bool = True) -> Dict[str, Tensor]:
    """Batch dict as produced by Data_Set.prepare_hybrid2_sample + default collate
    (data_loader/data_set.py:357-384): f32 images (B,3,H,W), f64 angles, int64 jitters.
    structured=True: per-sample low-frequency field + noise, view 2 correlated with view 1
    (white noise at default init is numerically chaotic)."""
    g = torch.Generator().manual_seed(seed)
    b = batch_size
    if structured:
        low = torch.randn(b, 3, 7, 7, generator=g)
        field = F.interpolate(low, size=(size, size), mode="bilinear", align_corners=False) * 1.5
        img1 = field + 0.25 * torch.randn(b, 3, size, size, generator=g)
        shift = size // 8
        img2 = 0.9 * torch.roll(field, shifts=(shift, -shift), dims=(2, 3)) + 0.25 * torch.randn(
            b, 3, size, size, generator=g)
    else:
        img1 = torch.randn(b, 3, size, size, generator=g)
        img2 = torch.randn(b, 3, size, size, generator=g)
    batch = {"transformed_image1": img1.contiguous(), "transformed_image2": img2.contiguous()}
    for k in (1, 2):
        batch[f"angle_{k}"] = torch.floor(torch.rand(b, generator=g, dtype=torch.float64) * 90 - 45)
        batch[f"jitter_x_{k}"] = -torch.randint(0, 15, (b,), generator=g, dtype=torch.int64)
        batch[f"jitter_y_{k}"] = -torch.randint(0, 15, (b,), generator=g, dtype=torch.int64)
    return batch


def oracle_step(model: OracleHybrid2Model, batch, optimizer=None, scheduler=None):
    """training_step -> backward -> optimizer step (what Lightning does per batch with
    accumulate_grad_batches == 1).  Returns the metric dict (detached)."""
    model.train()
    model.train_metrics = {}
    out = model.training_step(batch, 0)
    if optimizer is not None:
        optimizer.zero_grad()
    out["loss"].backward()
    if optimizer is not None:
        optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return {k: v.detach() for k, v in out.items()}


def named_grads(model: nn.Module) -> "OrderedDict[str, Tensor]":
    return OrderedDict((n, p.grad.detach().clone()) for n, p in model.named_parameters() if p.grad is not None)
