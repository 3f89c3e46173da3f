"""Checkpoint export: PeCLR ``encoder.features.*`` -> torchvision ResNet (contract of the reference's
src/models/port_model.py:7-48: positional copy with a per-entry name-suffix check)."""
import torch
import torchvision


def peclr_to_torchvision(resnet_model, path_to_peclr_weights):
    ckpt = torch.load(path_to_peclr_weights, map_location=torch.device("cpu"))
    state = ckpt["state_dict"]
    if not isinstance(resnet_model, torchvision.models.ResNet):
        raise Exception("The selected model is not of type ResNet from torch vision!")
    target = resnet_model.state_dict()
    target_items = list(target.items())
    feats = [(k, v) for k, v in state.items() if "features" in k]
    for idx, (key, value) in enumerate(feats):
        name, own = target_items[idx]
        if name.split(".")[-1] != key.split(".")[-1]:
            raise ValueError(f"PeCLR entry {key} does not line up with ResNet entry {name}")
        if own.shape != value.shape:
            raise ValueError(f"shape mismatch for {key} -> {name}: {tuple(value.shape)} vs {tuple(own.shape)}")
        own.copy_(value)
    return resnet_model
