"""Hot-path part of the reference's src/models/utils.py: encoder factory, the op-level loss-chain functions and
checkpoint path helpers.

Inside the training step the loss-chain ops of that file (vanila_contrastive_loss :154-186, rotate_encoding
:301-321, translate_encodings :325-346, get_rotation_2D_matrix :271-298) are executed by ONE fused CUDA kernel
(csrc/ntxent.cu).  The functions below expose each of them on its own behind the reference's signature (same
in-place semantics, same detached statistics, same autograd behaviour), each backed by its own CUDA kernel
(csrc/ntxent.cu plain mode, csrc/equiv_ops.cu); there is no torch fallback.
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


def _points(encoding: torch.Tensor):
    if encoding.dim() != 3 or encoding.shape[-1] < 2:
        raise ValueError("encoding must be (batch, points, >=2), got %s" % (tuple(encoding.shape),))
    if encoding.dtype != torch.float32:
        raise TypeError("the encodings kernels are fp32 (the reference's rotate_encoding only works in fp32 too)")
    return encoding.shape


class _RotateFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, encoding, angle):
        n, m, d = encoding.shape
        out = encoding.detach().contiguous().clone()
        rot = torch.empty((n, 4), dtype=torch.float32, device=out.device)
        ops.rotate_encoding_(out, angle, rot)
        ctx.save_for_backward(rot)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (rot,) = ctx.saved_tensors
        g = grad_out.contiguous().clone()
        ops.rotate_encoding_bwd_(g, rot)
        return g, None


def get_rotation_2D_matrix(angle: torch.Tensor, center_x: torch.Tensor, center_y: torch.Tensor, scale) -> torch.Tensor:
    """src/models/utils.py:271-298: (n,3,2) fp32 transposed OpenCV rotation matrices for a batch of angles in
    degrees.  (The reference allocates the result on the CPU; here it lives on the inputs' CUDA device.)"""
    return ops.rotation_2d_matrix(angle, center_x, center_y, float(scale))


def rotate_encoding(encoding: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    """src/models/utils.py:301-321: rotates every sample's 2-D points about their (detached) mean by `angle`
    degrees; in place on `encoding[..., :2]` and returns `encoding`, like the reference."""
    _points(encoding)
    rotated = _RotateFunction.apply(encoding, angle)
    encoding[..., :2] = rotated[..., :2]
    return encoding


class _TranslateFunction(torch.autograd.Function):
    """The shift is built from detached statistics (or is a constant), so the gradient is the identity."""

    @staticmethod
    def forward(ctx, encoding, translate_x, translate_y, exact):
        out = encoding.detach().contiguous().clone()
        ops.translate_encodings_(out, translate_x, translate_y, exact)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out, None, None, None


def _translate(encoding, translate_x, translate_y, exact):
    _points(encoding)
    shifted = _TranslateFunction.apply(encoding, translate_x, translate_y, exact)
    encoding[..., :2] = shifted[..., :2]
    return encoding


def translate_encodings(encoding: torch.Tensor, translate_x: torch.Tensor, translate_y: torch.Tensor) -> torch.Tensor:
    """src/models/utils.py:325-346: x += tx * (max_x - min_x), y += ty * (max_y - min_y) per sample (range over the
    points, detached); in place, returns `encoding`."""
    return _translate(encoding, translate_x, translate_y, False)


def translate_encodings2(encoding: torch.Tensor, translate_x: torch.Tensor, translate_y: torch.Tensor) -> torch.Tensor:
    """src/models/utils.py:349-364: exact translation x += tx, y += ty; in place, returns `encoding`."""
    return _translate(encoding, translate_x, translate_y, True)


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
