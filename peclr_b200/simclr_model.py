"""SimCLR shell (mirror of src/models/unsupervised/simclr_model.py:10-76): projection head + step functions.
The arithmetic of the step runs in peclr_b200.engine on the CUDA kernels."""
from typing import Dict

import torch
from torch import Tensor, nn

from .base_model import BaseModel
from .easydict import EasyDict as edict
from .ops import STAT_NAMES


class _StepFunction(torch.autograd.Function):
    """Connects the engine's explicit forward/backward to autograd: ``loss.backward()`` runs the CUDA backward
    pass, which accumulates into the flat gradient buffer every parameter's ``.grad`` is a view of."""

    @staticmethod
    def forward(ctx, anchor, model, loss, g_p, head_ctx, trunk_ctx):
        ctx.model, ctx.g_p, ctx.head_ctx, ctx.trunk_ctx = model, g_p, head_ctx, trunk_ctx
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.model.engine
        g_p = ctx.g_p * grad_out  # 2B x 128 scalar scale (loss / accumulate_grad_batches etc.)
        d_enc = eng.backward_head(g_p, ctx.head_ctx)
        eng.backward_trunk(d_enc, ctx.trunk_ctx, after_stage=ctx.model._after_stage_hook)
        eng.attach_grads()
        ctx.g_p = ctx.head_ctx = ctx.trunk_ctx = None
        return None, None, None, None, None, None


class SimCLR(BaseModel):
    uses_equivariance = False

    def __init__(self, config: edict):
        super().__init__(config)
        self.projection_head = self.get_projection_head()
        self._anchor = None
        self._after_stage_hook = None
        self._bind_engine()

    def get_projection_head(self) -> nn.Sequential:
        cfg = self.config
        return nn.Sequential(
            nn.Linear(cfg.projection_head_input_dim, cfg.projection_head_hidden_dim, bias=True),
            nn.BatchNorm1d(cfg.projection_head_hidden_dim),
            nn.ReLU(),
            nn.Linear(cfg.projection_head_hidden_dim, cfg.output_dim, bias=False),
        )

    # ---- the step --------------------------------------------------------------------------------------
    def _projections(self, batch: Dict[str, Tensor]):
        """trunk -> head, plus the per-sample correction parameters of the batch (None when not used)."""
        eng = self.engine
        img1, img2 = batch["transformed_image1"], batch["transformed_image2"]
        training = self.training
        enc, trunk_ctx = eng.forward_trunk(img1.contiguous(), img2.contiguous(), training=training)
        p, head_ctx = eng.forward_head(enc, training=training)
        aug = self.config.augmentation if self.uses_equivariance else []
        crop, rotate = "crop" in aug, "rotate" in aug
        angle = jx = jy = None
        if crop:
            jx = torch.cat((batch["jitter_x_1"], batch["jitter_x_2"])).to(torch.int64)
            jy = torch.cat((batch["jitter_y_1"], batch["jitter_y_2"])).to(torch.int64)
        if rotate:
            angle = torch.cat((batch["angle_1"], batch["angle_2"])).to(torch.float64)
        return p, (angle, jx, jy, tuple(img1.shape[-2:]), crop, rotate), head_ctx, trunk_ctx

    def _forward(self, batch: Dict[str, Tensor], want_grad: bool):
        """trunk -> head -> fused loss kernel.  Returns loss [1], stats [16], g_p (dloss/dp or None), contexts."""
        p, corr, head_ctx, trunk_ctx = self._projections(batch)
        loss, stats, g_p = self.engine.forward_loss(p, *corr, temperature=0.5, want_grad=want_grad)
        return loss, stats, g_p, head_ctx, trunk_ctx

    def _run_step(self, batch: Dict[str, Tensor], want_stats: bool):
        want_grad = self.training and torch.is_grad_enabled()
        loss, stats, g_p, head_ctx, trunk_ctx = self._forward(batch, want_grad)
        if want_grad:
            if self._anchor is None or self._anchor.device != loss.device:
                self._anchor = torch.zeros((), device=loss.device, requires_grad=True)
            loss = _StepFunction.apply(self._anchor, self, loss, g_p, head_ctx, trunk_ctx)
        else:
            loss = loss.reshape(())
        return loss, stats

    def forward_backward(self, batch: Dict[str, Tensor], grad_scale: float = 1.0) -> Dict[str, Tensor]:
        """training_step + (loss * grad_scale).backward() without going through autograd: the same kernels in
        the same order, enqueued directly (this is what the CUDA-graph capture records).  Gradients accumulate
        into the flat buffer every parameter's .grad is a view of.  Returns the metric dict (detached)."""
        with torch.no_grad():
            loss, stats, g_p, head_ctx, trunk_ctx = self._forward(batch, want_grad=True)
            if grad_scale != 1.0:
                g_p = g_p * grad_scale
            d_enc = self.engine.backward_head(g_p, head_ctx)
            self.engine.backward_trunk(d_enc, trunk_ctx, after_stage=self._after_stage_hook)
        out = {"loss": loss.reshape(())}
        if self.uses_equivariance:
            out.update({name: stats[i] for i, name in enumerate(STAT_NAMES)})
        self.train_metrics = {**self.train_metrics, **out}
        return out

    def contrastive_step(self, batch: Dict[str, Tensor]) -> Tensor:
        loss, _ = self._run_step(batch, want_stats=False)
        return loss

    def get_encodings(self, batch_images: Tensor) -> Tensor:
        return self.encoder(batch_images)

    def forward(self, x: Tensor) -> Dict[str, Tensor]:
        embedding = self.encoder(x)
        projection, _ = self.engine.forward_head(embedding, training=self.training)
        return {"embedding": embedding, "projection": projection}

    def training_step(self, batch: dict, batch_idx: int) -> Dict[str, Tensor]:
        loss = self.contrastive_step(batch)
        self.train_metrics = {**self.train_metrics, **{"loss": loss}}
        self.plot_params = {
            "image1": batch["transformed_image1"],
            "image2": batch["transformed_image2"],
            "params": {k: v for k, v in batch.items() if "image" not in k},
        }
        return self.train_metrics

    def validation_step(self, batch: dict, batch_idx: int) -> Dict[str, Tensor]:
        loss = self.contrastive_step(batch)
        self.plot_params = {
            "image1": batch["transformed_image1"],
            "image2": batch["transformed_image2"],
            "params": {k: v for k, v in batch.items() if "image" not in k},
        }
        return {"loss": loss}
