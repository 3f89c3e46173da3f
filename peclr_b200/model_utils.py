"""Hot-path part of the reference's src/models/utils.py: encoder factory and checkpoint path helpers.

The loss-chain ops of that file (vanila_contrastive_loss :154-186, rotate_encoding :301-321,
translate_encodings :325-346, get_rotation_2D_matrix :271-298) are executed by the fused CUDA kernel
(csrc/ntxent.cu); `vanila_contrastive_loss` below exposes the NT-Xent part on its own, same signature.
"""
import os

import torch

from . import ops
from .easydict import EasyDict as edict
from .resnet_model import ResNetModel


def get_wrapper_model(config: edict, pretrained: bool, wrapper: bool = False):
    """src/models/utils.py:412-428 (the `wrapper=True` branch of the reference names an undefined class)."""
    cfg = edict({
        "model": {"backend_model": "resnet" + config.resnet_size, "norm_layer": "bn", "use_var": False,
                  "pretrained": pretrained},
        "dataset": {"np": 21},
        "loss": {"hmap": {"enabled": False}},
    })
    if wrapper:
        raise NotImplementedError("WrapperModel is undefined in the reference as well")
    return ResNetModel(config=cfg, mode="pretraining")


class _NtXentFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, temperature):
        z = torch.cat([z1, z2], dim=0).contiguous().float()
        loss, _, g = ops.ntxent_plain(z, temperature)
        ctx.save_for_backward(g)
        ctx.b = z1.shape[0]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        g = g * grad_out
        return g[: ctx.b], g[ctx.b:], None


def vanila_contrastive_loss(z1: torch.Tensor, z2: torch.Tensor, temperature: float = 0.5) -> torch.Tensor:
    """NT-Xent of SimCLR over the 2N batch (self excluded, positive included), computed by the fused CUDA kernel
    in its plain mode (no normalisation / equivariance correction).  Same contract as the reference function."""
    return _NtXentFunction.apply(z1, z2, temperature)


def get_latest_checkpoint(experiment_name: str, checkpoint: str = "") -> str:
    """src/models/utils.py:189-206: $SAVED_MODELS_BASE_PATH/<experiment>/checkpoints/epoch=<int>.ckpt"""
    base = os.environ.get("SAVED_MODELS_BASE_PATH", "")
    path = os.path.join(base, experiment_name, "checkpoints")
    if checkpoint == "":
        checkpoint = sorted(os.listdir(path), key=lambda x: int(x[6:-5]))[-1]
    return os.path.join(path, checkpoint)


def get_encoder_state_dict(saved_model_path: str, checkpoint: str) -> dict:
    """src/models/utils.py:209-225: encoder.* entries with the 8-character "encoder." prefix removed."""
    sd = torch.load(get_latest_checkpoint(saved_model_path, checkpoint), map_location="cpu")["state_dict"]
    return {k[8:]: v for k, v in sd.items() if "encoder" in k}
