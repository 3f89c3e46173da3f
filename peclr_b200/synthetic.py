"""Synthetic two-view generator that stands in for the reference's src/data_loader (FreiHAND / YT3DH datasets +
OpenCV SampleAugmenter, out of scope here).  It emits the exact batch-dict schema of
Data_Set.prepare_hybrid2_sample + default collate (src/data_loader/data_set.py:357-384):

    transformed_image1/2  float32 (B,3,H,W), ImageNet-normalised statistics (mean 0, std 1)
    angle_1/2             float64 (B,), integer valued in [-45, 45)
    jitter_x_1/2, jitter_y_1/2   int64 (B,), in {-14..0}
"""
from typing import Dict

import torch
import torch.nn.functional as F
from torch.utils.data import Dataset


def synthetic_batch(batch_size: int, size: int, seed: int = 5, structured: bool = True,
                    pin_memory: bool = False) -> Dict[str, torch.Tensor]:
    """One two-view batch on the host.  structured=True gives per-sample low-frequency fields + noise with view 2
    correlated to view 1 (what parity runs use: white noise at default init is numerically chaotic,
    SURVEY 3.6); structured=False gives plain N(0,1) images (throughput runs)."""
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
    if pin_memory:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


class SyntheticTwoViewDataset(Dataset):
    """Per-sample version for a DataLoader: item i is deterministic in (seed, i)."""

    def __init__(self, num_samples: int, size: int, seed: int = 5, structured: bool = True, rotate: bool = True):
        self.num_samples, self.size, self.seed, self.structured, self.rotate = num_samples, size, seed, structured, rotate

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        b = synthetic_batch(1, self.size, seed=self.seed * 1000003 + idx, structured=self.structured)
        out = {}
        for k, v in b.items():
            if "angle" in k and not self.rotate:
                continue  # prepare_hybrid2_sample drops None entries (data_set.py:382-383)
            out[k] = v[0] if v.dim() > 1 else (float(v[0]) if v.dtype == torch.float64 else int(v[0]))
        return out
