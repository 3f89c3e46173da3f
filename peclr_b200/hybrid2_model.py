"""PeCLR model (mirror of src/models/unsupervised/hybrid2_model.py:16-106): SimCLR whose projections are
moved back through the inverse crop-translation and inverse rotation before NT-Xent.  The whole chain after
the projection head -- statistics, both normalisations, translate, rotate, NT-Xent and its backward -- is one
CUDA launch (csrc/ntxent.cu)."""
from typing import Dict, Tuple

import torch
from torch import Tensor

from . import ops
from .easydict import EasyDict as edict
from .ops import STAT_NAMES
from .simclr_model import SimCLR


class Hybrid2Model(SimCLR):
    uses_equivariance = True

    def __init__(self, config: edict):
        super().__init__(config)

    def contrastive_step(self, batch: Dict[str, Tensor]) -> Tensor:
        loss, stats = self._run_step(batch, want_stats=True)
        # the 16 detached projection statistics of get_projection_stats (hybrid2_model.py:92-106)
        self.train_metrics = {**self.train_metrics, **{name: stats[i] for i, name in enumerate(STAT_NAMES)}}
        return loss

    def get_transformed_projections(self, batch: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
        """hybrid2_model.py:27-85: the two views' projections after normalise -> un-translate -> un-rotate ->
        normalise (the z rows the fused kernel feeds to NT-Xent), plus the 16 projection statistics merged into
        ``train_metrics``.  Inspection API: the result is detached -- training goes through ``training_step``,
        whose backward runs inside the same fused launch."""
        with torch.no_grad():
            p, corr, _, _ = self._projections(batch)
            b = p.shape[0] // 2
            ws = ops.ntxent_workspace(b, 1, p.device)  # private workspace: local rows only, also under DDP
            _, stats, _ = ops.ntxent_fused(p, *corr, temperature=0.5, want_grad=False, workspace=ws)
        z = ws[: 2 * b * p.shape[1]].view(2 * b, p.shape[1])
        self.train_metrics = {**self.train_metrics, **{name: stats[i] for i, name in enumerate(STAT_NAMES)}}
        return z[:b].clone(), z[b:].clone()

    def get_projection_stats(self, projection: Tensor, name: str) -> dict:
        """hybrid2_model.py:92-106 for a (batch, points, 2) tensor: batch means of the per-sample mean / (lower)
        median / min / max of each coordinate, computed by one CUDA launch (csrc/equiv_ops.cu)."""
        st = ops.projection_stats(projection.detach().float().contiguous())
        keys = [f"{name}{c}_{s}" for c in "xy" for s in ("mean", "median", "min", "max")]
        return {k: st[i] for i, k in enumerate(keys)}
