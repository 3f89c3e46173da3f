"""Optimiser + schedule of the reference's BaseModel.configure_optimizers (src/models/base_model.py:57-104),
backed by the fused multi-tensor CUDA kernel (csrc/lars_adam.cu).

`FusedLARSAdam` plays the role of ``LARSWrapper(torch.optim.Adam(groups, lr))`` (pl_bolts 0.2.2, eta=0.02,
clip=True, eps=1e-8) or of plain ``torch.optim.Adam`` (lars=False); `LinearWarmupCosineAnnealingLR` restates
pl_bolts' scheduler of the same name (recursive form, including its warmup_epochs == 0 quirk).
"""
import math

import torch
from torch.optim.lr_scheduler import LRScheduler


class FusedLARSAdam(torch.optim.Optimizer):
    def __init__(self, params, engine, lr, lars=True, betas=(0.9, 0.999), eps=1e-8, eta=0.02, clip=True,
                 lars_eps=1e-8):
        super().__init__(params, dict(lr=lr, weight_decay=0.0, betas=betas, eps=eps))
        self.engine = engine
        self.lars, self.eta, self.clip, self.lars_eps = lars, eta, clip, lars_eps
        self.step_count = 0
        self._tables = None
        # a new optimiser starts from fresh Adam moments (they live in the engine's flat buffers; load_state_dict
        # puts a checkpoint's moments back)
        engine.exp_avg = engine.exp_avg_sq = None

    def _seg_weight_decay(self):
        wd_of = {}
        for g in self.param_groups:
            for p in g["params"]:
                wd_of[id(p)] = g["weight_decay"]
        out = []
        for s in self.engine.segs:
            p = s.module._parameters[s.pname]
            if id(p) not in wd_of:
                raise ValueError(f"parameter {s.name} is not in any optimiser group")
            out.append(wd_of[id(p)])
        return out

    @torch.no_grad()
    def step(self, closure=None):
        from . import ops

        eng = self.engine
        eng._require_cuda()
        lrs = {g["lr"] for g in self.param_groups}
        if len(lrs) != 1:
            raise NotImplementedError("FusedLARSAdam expects one learning rate shared by all groups")
        wds = self._seg_weight_decay()
        if self._tables is None or self._tables["wd_list"] != wds:
            self._tables = ops.build_opt_tables([s.size for s in eng.segs], wds, eng.device)
            self._tables["wd_list"] = wds
        if eng.exp_avg is None:
            eng.exp_avg = torch.zeros_like(eng.flat)
            eng.exp_avg_sq = torch.zeros_like(eng.flat)
        if eng.w_bf16 is None:
            eng.sync_weights()
        self.step_count += 1
        g0 = self.param_groups[0]
        ops.lars_adam_step(eng.flat, eng.grads, eng.exp_avg, eng.exp_avg_sq, self._tables, lrs.pop(), self.step_count,
                           p_bf16=eng.w_bf16, lars=self.lars, betas=g0["betas"], adam_eps=g0["eps"], eta=self.eta,
                           clip=self.clip, lars_eps=self.lars_eps)
        eng.refresh_derived_weights()
        eng.weights_dirty = False

    def zero_grad(self, set_to_none: bool = False):
        # gradients live in the engine's flat buffer; parameters keep their .grad views
        self.engine.zero_grad()

    def state_dict(self):
        sd = super().state_dict()
        eng = self.engine
        sd["fused"] = dict(step_count=self.step_count,
                           exp_avg=None if eng.exp_avg is None else eng.exp_avg.detach().cpu(),
                           exp_avg_sq=None if eng.exp_avg_sq is None else eng.exp_avg_sq.detach().cpu())
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        fused = sd.pop("fused", None)
        super().load_state_dict(sd)
        if fused is not None:
            self.step_count = fused["step_count"]
            if fused["exp_avg"] is not None:
                self.engine.exp_avg = fused["exp_avg"].to(self.engine.device)
                self.engine.exp_avg_sq = fused["exp_avg_sq"].to(self.engine.device)


class LinearWarmupCosineAnnealingLR(LRScheduler):
    """Linear warm-up from `warmup_start_lr` to the base lr over `warmup_epochs` scheduler steps, then cosine
    annealing to `eta_min` at `max_epochs` (pl_bolts 0.2.2 semantics; "epochs" are scheduler steps)."""

    def __init__(self, optimizer, warmup_epochs, max_epochs, warmup_start_lr=0.0, eta_min=0.0, last_epoch=-1):
        self.warmup_epochs, self.max_epochs = warmup_epochs, max_epochs
        self.warmup_start_lr, self.eta_min = warmup_start_lr, eta_min
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
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
        num = 1 + math.cos(math.pi * (e - w) / (m - w))
        den = 1 + math.cos(math.pi * (e - w - 1) / (m - w))
        return [num / den * (g["lr"] - self.eta_min) + self.eta_min for g in groups]
