"""LightningModule-shaped base of the PeCLR models (mirror of src/models/base_model.py:13-127): encoder
construction, weight-decay exclusion groups, LARS-Adam + schedule, epoch-end metric reduction."""
import math
from typing import Dict, Iterator, List, Tuple, Union

import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from .easydict import EasyDict as edict
from .lightning import LightningModule
from .model_utils import get_wrapper_model
from .optim import FusedLARSAdam, LinearWarmupCosineAnnealingLR


class BaseModel(LightningModule):
    def __init__(self, config: edict):
        super().__init__()
        if "resnet_size" in config.keys():
            self.encoder = get_wrapper_model(config, pretrained=True)
        self.config = config
        self.train_metrics_epoch = {}
        self.train_metrics = {}
        self.validation_metrics_epoch = {}
        self.plot_params = {}
        self.engine = None

    # ---- engine binding: flat parameter arena + kernels -------------------------------------------------
    def _bind_engine(self):
        from .engine import StepEngine

        self.engine = StepEngine(self)
        self.encoder.engine = self.engine

    def _apply(self, fn, recurse=True):
        """Device moves re-home the flat buffers instead of scattering the parameter views."""
        if self.engine is None:
            return super()._apply(fn, recurse)
        probe = fn(torch.empty(0, dtype=torch.float32, device=self.engine.device))
        if probe.dtype != torch.float32:
            raise TypeError("peclr_b200 keeps fp32 master weights; the compute precision is fixed by the kernels")
        self.engine.to(probe.device)
        return self

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict)
        if self.engine is not None:
            self.engine.weights_dirty = True
        return out

    def zero_grad(self, set_to_none: bool = False):
        """Gradients live in the engine's flat buffer; every parameter's .grad stays a view of it."""
        if self.engine is None:
            return super().zero_grad(set_to_none)
        self.engine.zero_grad()
        self.engine.attach_grads()

    # ---- data-parallel gradient exchange (SUM over ranks: the fused loss already returns the gradient of the
    # GLOBAL-batch mean w.r.t. the local rows) -------------------------------------------------------
    # Overlapped with backward: the flat gradient buffer is in named_parameters() order, so the ResNet stages are four
    # contiguous ranges [stem + layer1 | layer2 | layer3 | layer4 + head].  backward_trunk reports a stage as soon as
    # its gradients are final (layer4 first); its range is all-reduced right then on a communication stream, next to
    # the rest of the backward pass.  Inside a CUDA-graph capture the collectives are captured with the step.
    @staticmethod
    def _dp_world():
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _stage_ranges(self):
        eng = self.engine
        if getattr(self, "_ranges_for", None) is not eng.segs:
            first = {}
            for s in eng.segs:
                key = s.name.split(".")[2] if s.name.startswith("encoder.features.") else "head"
                first.setdefault(key, s.begin)
            l2, l3, l4 = first["5"], first["6"], first["7"]
            # reported stage -> range to reduce: 3 = layer4 (+ head, final since backward_head), 2 = layer3,
            # 1 = layer2, -1 = end of backward (layer1 + stem)
            self._ranges = {3: (l4, eng.total), 2: (l3, l4), 1: (l2, l3), -1: (0, l2)}
            self._ranges_for = eng.segs
        return self._ranges

    def enable_overlapped_sync(self, on: bool = True):
        """Arms the per-stage all-reduce for the backward passes that follow (the Trainer / bench arm it for the micro-step
        that closes an accumulation window).  No-op on one GPU."""
        if on and self._dp_world() > 1:
            self._after_stage_hook = self._reduce_stage
        else:
            self._after_stage_hook = None

    def _reduce_stage(self, stage):
        import torch.distributed as dist

        rng = self._stage_ranges().get(stage)
        if rng is None:
            return
        eng = self.engine
        if getattr(self, "_comm", None) is None or self._comm.device != eng.device:
            self._comm = torch.cuda.Stream(device=eng.device)
        cur = torch.cuda.current_stream()
        self._comm.wait_stream(cur)
        if eng.overlap_wgrad and eng._side is not None:
            self._comm.wait_stream(eng._side)  # the stage's weight gradients run on the engine's side stream
        with torch.cuda.stream(self._comm):
            dist.all_reduce(eng.grads[rng[0]:rng[1]], op=dist.ReduceOp.SUM)
        self._pending_sync = True
        if stage == -1:
            cur.wait_stream(self._comm)  # end of backward: join (inside a capture this closes the forked stream)

    def sync_gradients(self):
        """Completes the data-parallel gradient exchange of this optimiser step: waits for the per-stage reductions
        if the backward pass issued them, else reduces the whole flat buffer in one call."""
        import torch.distributed as dist

        if self._dp_world() <= 1:
            return
        if getattr(self, "_pending_sync", False):
            torch.cuda.current_stream().wait_stream(self._comm)
            self._pending_sync = False
        else:
            dist.all_reduce(self.engine.grads, op=dist.ReduceOp.SUM)

    # ---- reference API -----------------------------------------------------------------------------------
    def exclude_from_wt_decay(
        self,
        named_params: Iterator[Tuple[str, torch.Tensor]],
        weight_decay: float,
        skip_list: List[str] = ["bias", "bn"],
    ) -> List[Dict[str, Union[list, float]]]:
        decayed, excluded = [], []
        for name, param in named_params:
            if not param.requires_grad:
                continue
            (excluded if any(key in name for key in skip_list) else decayed).append(param)
        return [{"params": decayed, "weight_decay": weight_decay}, {"params": excluded, "weight_decay": 0.0}]

    def setup(self, stage: str):
        global_batch_size = self.trainer.world_size * self.config.batch_size
        self.train_iters_per_epoch = self.config.num_samples // global_batch_size

    def configure_optimizers(self) -> Tuple[list, list]:
        cfg = self.config
        # encoder.final_layer never receives a gradient in pre-training (torch.optim.Adam skips it in the
        # reference, base_model.py:62); it is not part of the fused optimiser's flat buffer.
        # it stays in the param groups (as in the reference) but not in the fused optimiser's flat buffer.
        groups = self.exclude_from_wt_decay(self.named_parameters(), weight_decay=cfg.opt_weight_decay)
        lr = cfg.lr * math.sqrt(cfg.batch_size * cfg.num_of_mini_batch)
        warmup = cfg.warmup_epochs * self.train_iters_per_epoch // cfg.num_of_mini_batch
        if "lr_max_epochs" in cfg.keys() and cfg["lr_max_epochs"] is not None:
            max_steps = cfg["lr_max_epochs"] * self.train_iters_per_epoch // cfg.num_of_mini_batch
        else:
            max_steps = self.trainer.max_epochs * self.train_iters_per_epoch // cfg.num_of_mini_batch
        if cfg.optimizer == "LARS":
            optimizer = FusedLARSAdam(groups, self.engine, lr=lr, lars=True)
            scheduler = LinearWarmupCosineAnnealingLR(optimizer, warmup_epochs=warmup, max_epochs=max_steps,
                                                      warmup_start_lr=0, eta_min=0)
        else:
            optimizer = FusedLARSAdam(groups, self.engine, lr=lr, lars=False)
            scheduler = CosineAnnealingLR(optimizer, T_max=max_steps)
        return [optimizer], [{"scheduler": scheduler, "interval": "step", "frequency": 1}]

    def training_epoch_end(self, outputs: List[dict]):
        keys = outputs[0].keys()
        self.train_metrics_epoch = {k: torch.stack([o[k] for o in outputs]).mean() for k in keys}
        monitored = "loss_3d" if "loss_3d" in keys else "loss"
        self.log("checkpoint_saving_loss", self.train_metrics_epoch[monitored])

    def validation_epoch_end(self, outputs: List[dict]):
        keys = outputs[0].keys()
        self.validation_metrics_epoch = {k: torch.stack([o[k] for o in outputs]).mean() for k in keys}
