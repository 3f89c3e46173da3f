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

    def sync_gradients(self):
        """Data-parallel gradient exchange (sum over ranks; see peclr_b200.lightning)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
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
