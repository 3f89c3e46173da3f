"""Minimal stand-ins for the parts of pytorch-lightning 1.0.8 the reference's pre-training path uses
(pytorch_lightning is not installed): a LightningModule base, a Trainer that runs the same hook sequence
(setup -> configure_optimizers -> per batch training_step / backward / optimizer + scheduler step every
`accumulate_grad_batches` batches -> training_epoch_end -> validation -> checkpoint), a top-k
ModelCheckpoint writing ``epoch=N.ckpt`` files with the ``{"state_dict": ...}`` layout
(src/models/callbacks/model_checkpoint.py:5-10, src/experiments/peclr_training.py:73-96), and
seed_everything.  Data-parallel runs launch one process per GPU (torchrun); gradients are SUMMED across
ranks because the fused loss already returns the gradient of the GLOBAL-batch mean.
"""
import json
import os
import random
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn


def seed_everything(seed: int) -> int:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    return seed


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = None
        self._logged: Dict[str, float] = {}

    def log(self, name, value, **kwargs):
        self._logged[name] = value

    # hooks with the reference's names; subclasses override what they need
    def setup(self, stage: str):
        pass

    def training_epoch_end(self, outputs):
        pass

    def validation_epoch_end(self, outputs):
        pass


class Callback:
    def on_train_batch_end(self, trainer, module, outputs, batch, batch_idx):
        pass

    def on_epoch_end(self, trainer, module):
        pass


class LearningRateMonitor(Callback):
    def __init__(self, logging_interval="epoch"):
        self.logging_interval = logging_interval
        self.lrs: List[float] = []

    def on_train_batch_end(self, trainer, module, outputs, batch, batch_idx):
        if self.logging_interval == "step":
            self.lrs.append(trainer.optimizers[0].param_groups[0]["lr"])

    def on_epoch_end(self, trainer, module):
        if self.logging_interval == "epoch":
            self.lrs.append(trainer.optimizers[0].param_groups[0]["lr"])


class ModelCheckpoint(Callback):
    """Top-k checkpoints by a monitored metric (lower is better), files ``<dirpath>/epoch=N.ckpt``."""

    def __init__(self, save_top_k=1, period=1, monitor="checkpoint_saving_loss", dirpath=None):
        self.save_top_k, self.period, self.monitor, self.dirpath = save_top_k, period, monitor, dirpath
        self.best: List[tuple] = []  # (value, path)
        self._pending = None

    def _save_model(self, filepath: str, trainer, pl_module):
        print(f"Saving checkpoint at {filepath}")  # (UpdatedModelCheckpoint, callbacks/model_checkpoint.py:5-10)
        os.makedirs(os.path.dirname(filepath), exist_ok=True)
        ckpt = {
            "epoch": trainer.current_epoch,
            "global_step": trainer.global_step,
            "state_dict": pl_module.state_dict(),
            "optimizer_states": [o.state_dict() for o in trainer.optimizers],
            "lr_schedulers": [s["scheduler"].state_dict() for s in trainer.lr_schedulers],
            # top-k bookkeeping (so that a resumed run keeps pruning the files of the run it continues) and where the
            # encoder's initial weights came from (ImageNet file or random: see ResNetModel._init_like_reference)
            "callbacks": {"ModelCheckpoint": {"best": [(v, p) for v, p in self.best] + [(self._pending, filepath)],
                                              "monitor": self.monitor}},
            "init_source": getattr(getattr(pl_module, "encoder", None), "init_source", None),
        }
        torch.save(ckpt, filepath)

    def restore_state(self, ckpt: dict, ckpt_path: str):
        """Rebuilds the top-k list on resume: from the checkpoint's own record, else (a checkpoint written by the
        reference) from the ``epoch=N.ckpt`` files next to it, with an unknown (infinite) metric: they are the
        first to be pruned once better-ranked checkpoints exist."""
        rec = (ckpt.get("callbacks") or {}).get("ModelCheckpoint")
        if rec and rec.get("best"):
            best = [(float(v), p) for v, p in rec["best"] if os.path.exists(p)]
        else:
            d = os.path.dirname(os.path.abspath(ckpt_path))
            found = []
            for name in os.listdir(d):
                if name.startswith("epoch=") and name.endswith(".ckpt"):
                    try:
                        found.append((int(name[len("epoch="):-len(".ckpt")]), os.path.join(d, name)))
                    except ValueError:
                        pass
            best = [(float("inf"), p) for _, p in sorted(found)]
        self.best = sorted(best, key=lambda t: t[0])

    def on_epoch_end(self, trainer, module):
        if trainer.global_rank != 0 or self.save_top_k == 0 or (trainer.current_epoch + 1) % self.period:
            return
        value = module._logged.get(self.monitor)
        if value is None:
            return
        value = float(value)
        dirpath = self.dirpath or os.path.join(trainer.default_root_dir, "checkpoints")
        path = os.path.join(dirpath, f"epoch={trainer.current_epoch}.ckpt")
        if self.save_top_k < 0 or len(self.best) < self.save_top_k or value < max(v for v, _ in self.best):
            self._pending = value
            self._save_model(path, trainer, module)
            self.best.append((value, path))
            self.best.sort(key=lambda t: t[0])
            while 0 < self.save_top_k < len(self.best):
                _, drop = self.best.pop()
                if os.path.exists(drop):
                    os.remove(drop)


class JsonlLogger:
    """Stdout / JSONL metric sink replacing the Comet logger (no network; optional in the reference's README)."""

    def __init__(self, save_dir=None, experiment_name="peclr"):
        self.save_dir, self.experiment_name = save_dir, experiment_name
        self._fh = None
        if save_dir:
            os.makedirs(save_dir, exist_ok=True)
            self._fh = open(os.path.join(save_dir, experiment_name + ".jsonl"), "a")

    def get_key(self) -> str:
        """Stable identifier of this run (the role of comet's experiment key in save_experiment_key)."""
        import hashlib

        return hashlib.sha1(f"{self.save_dir}|{self.experiment_name}".encode()).hexdigest()[:32]

    def log_metrics(self, metrics: dict, step: Optional[int] = None):
        rec = {"step": step, **{k: float(v) for k, v in metrics.items()}}
        if self._fh:
            self._fh.write(json.dumps(rec) + "\n")
            self._fh.flush()

    def log_hyperparams(self, params):
        if self._fh:
            self._fh.write(json.dumps({"hparams": {k: str(v) for k, v in dict(params).items()}}) + "\n")


class Trainer:
    def __init__(self, accumulate_grad_batches=1, gpus=None, logger=None, max_epochs=1, precision=16,
                 amp_backend="native", callbacks=None, checkpoint_callback=None, default_root_dir=None,
                 limit_train_batches=None, limit_val_batches=None, log_every_n_steps=50, use_cuda_graph=True,
                 resume_from_checkpoint=None):
        self.accumulate_grad_batches = accumulate_grad_batches
        self.max_epochs = max_epochs
        self.logger = logger
        self.callbacks = list(callbacks or [])
        self.checkpoint_callback = checkpoint_callback
        if checkpoint_callback not in (None, False, True):
            self.callbacks.append(checkpoint_callback)
        self.default_root_dir = default_root_dir or os.environ.get("SAVED_META_INFO_PATH") or os.getcwd()
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.log_every_n_steps = log_every_n_steps
        self.use_cuda_graph = use_cuda_graph
        self.resume_from_checkpoint = resume_from_checkpoint  # path of an ``epoch=N.ckpt`` written by ModelCheckpoint
        # precision: the trunk computes in bf16 on the tensor cores with fp32 accumulation / master weights; the
        # reference's fp16 AMP + GradScaler (precision=16) has no counterpart to configure here.
        self.precision = precision
        import torch.distributed as dist

        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.global_rank = dist.get_rank() if self.world_size > 1 else 0
        self.current_epoch = 0
        self.global_step = 0
        self.optimizers, self.lr_schedulers = [], []
        self.images_per_sec = None

    @staticmethod
    def _with_last_flag(loader, limit):
        """(index, batch, is_last_batch_of_the_epoch) with one batch of look-ahead."""
        it = iter(loader)
        try:
            cur = next(it)
        except StopIteration:
            return
        idx = 0
        while True:
            try:
                nxt = next(it) if (limit is None or idx + 1 < limit) else None
            except StopIteration:
                nxt = None
            yield idx, cur, nxt is None
            if nxt is None:
                return
            cur, idx = nxt, idx + 1

    def _to_device(self, batch, device):
        return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def restore(self, path: str, model: LightningModule) -> int:
        """Loads a checkpoint written by ModelCheckpoint (Lightning's layout: state_dict, optimizer_states,
        lr_schedulers, epoch, global_step) into the model and this trainer's optimiser / scheduler.  Returns the epoch
        to continue with.  A checkpoint without optimiser state (e.g. written by the reference) restores weights only."""
        ckpt = torch.load(path, map_location="cpu")
        model.load_state_dict(ckpt["state_dict"])
        for opt, sd in zip(self.optimizers, ckpt.get("optimizer_states", [])):
            opt.load_state_dict(sd)
        for sch, sd in zip(self.lr_schedulers, ckpt.get("lr_schedulers", [])):
            sch["scheduler"].load_state_dict(sd)
        self.global_step = int(ckpt.get("global_step", 0))
        for cb in self.callbacks:
            if isinstance(cb, ModelCheckpoint):
                cb.restore_state(ckpt, path)
        return int(ckpt.get("epoch", -1)) + 1

    def fit(self, model: LightningModule, train_dataloader, val_dataloader=None):
        device = torch.device("cuda", torch.cuda.current_device())
        model.trainer = self
        model.to(device)
        if self.world_size > 1:
            model.engine.world, model.engine.rank = self.world_size, self.global_rank
        model.setup("fit")
        init_source = getattr(getattr(model, "encoder", None), "init_source", None)
        if init_source is not None and self.global_rank == 0:
            print(f"encoder initial weights: {init_source}")
            if self.logger:
                self.logger.log_hyperparams({"encoder_init_source": init_source})
        opts, scheds = model.configure_optimizers()
        self.optimizers, self.lr_schedulers = opts, scheds
        opt, sched = opts[0], scheds[0]["scheduler"]
        acc = self.accumulate_grad_batches
        graphed = None
        first_epoch = self.restore(self.resume_from_checkpoint, model) if self.resume_from_checkpoint else 0
        opt.zero_grad()
        for epoch in range(first_epoch, self.max_epochs):
            self.current_epoch = epoch
            model.train()
            outputs = []
            t0, seen = time.time(), 0
            for batch_idx, batch, is_last in self._with_last_flag(train_dataloader, self.limit_train_batches):
                batch = self._to_device(batch, device)
                # Lightning 1.0.8 steps when the accumulation window is full AND on the final batch of the epoch (a
                # partial window: the loss keeps its 1 / accumulate_grad_batches scale), so no gradient leaks into
                # the next epoch's first window
                closing = (batch_idx + 1) % acc == 0 or is_last
                synced = False
                # data parallel: the micro-step that closes the window all-reduces its gradients stage by stage
                # while its backward pass is still running
                model.enable_overlapped_sync(closing)
                if self.use_cuda_graph:
                    # same work as the eager branch below, submitted as one captured CUDA graph
                    if graphed is None:
                        from .graphed import GraphedStep

                        graphed = GraphedStep(model, batch, grad_scale=1.0 / acc)
                        opt.zero_grad()
                    if graphed.matches(batch):
                        out = graphed(batch, sync=closing)
                        synced = graphed.synced
                        outputs.append({k: v.clone() for k, v in out.items()})
                    else:  # e.g. a short last batch: same kernels, submitted eagerly
                        out = model.forward_backward(batch, 1.0 / acc)
                        outputs.append({k: v.detach() for k, v in out.items()})
                else:
                    model.train_metrics = {}
                    out = model.training_step(batch, batch_idx)
                    (out["loss"] / acc).backward()
                    outputs.append({k: v.detach() for k, v in out.items()})
                seen += 2 * batch["transformed_image1"].shape[0] * self.world_size
                if closing:
                    if not synced:
                        model.sync_gradients()
                    opt.step()
                    opt.zero_grad()
                    sched.step()
                    self.global_step += 1
                for cb in self.callbacks:
                    cb.on_train_batch_end(self, model, out, batch, batch_idx)
                if self.logger and self.global_rank == 0 and batch_idx % self.log_every_n_steps == 0:
                    self.logger.log_metrics({k: v for k, v in outputs[-1].items()}, step=self.global_step)
            torch.cuda.synchronize()
            self.images_per_sec = seen / max(time.time() - t0, 1e-9)
            if outputs:
                model.training_epoch_end(outputs)
            if val_dataloader is not None:
                model.eval()
                vouts = []
                with torch.no_grad():
                    for batch_idx, batch in enumerate(val_dataloader):
                        if self.limit_val_batches is not None and batch_idx >= self.limit_val_batches:
                            break
                        vouts.append(model.validation_step(self._to_device(batch, device), batch_idx))
                if vouts:
                    model.validation_epoch_end(vouts)
            for cb in self.callbacks:
                cb.on_epoch_end(self, model)
        return model
