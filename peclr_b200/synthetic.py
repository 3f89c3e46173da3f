"""Synthetic two-view generator that stands in for the reference's src/data_loader (FreiHAND / YT3DH datasets +
OpenCV SampleAugmenter, out of scope here).  It emits the exact batch-dict schema of
Data_Set.prepare_hybrid2_sample + default collate (src/data_loader/data_set.py:357-384):

    transformed_image1/2  float32 (B,3,H,W), ImageNet-normalised statistics (mean 0, std 1)
    angle_1/2             float64 (B,), integer valued in [-45, 45)
    jitter_x_1/2, jitter_y_1/2   int64 (B,), in {-14..0}
"""
import copy
from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset


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

    def __init__(self, num_samples: int, size: int, seed: int = 5, structured: bool = True, rotate: bool = True,
                 train_ratio: float = 1.0):
        self.num_samples, self.size, self.seed, self.structured, self.rotate = num_samples, size, seed, structured, rotate
        # train / validation split with the reference's interface (Data_Set.is_training, data_set.py:386-398; the
        # FreiHAND loader draws the split with train_test_split(train_size=train_ratio, random_state=seed),
        # freihand_loader.py:46-60).  Synthetic samples are i.i.d. in their index, so a contiguous split is
        # equivalent: the first round(train_ratio * N) indices train, the rest validate.
        self.num_train = min(num_samples, int(round(train_ratio * num_samples)))
        self._training = True

    def is_training(self, value: bool):
        self._training = bool(value)

    def __len__(self):
        return self.num_train if self._training else self.num_samples - self.num_train

    def __getitem__(self, idx):
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        if not self._training:
            idx += self.num_train
        b = synthetic_batch(1, self.size, seed=self.seed * 1000003 + idx, structured=self.structured)
        out = {}
        for k, v in b.items():
            if "angle" in k and not self.rotate:
                continue  # prepare_hybrid2_sample drops None entries (data_set.py:382-383)
            out[k] = v[0] if v.dim() > 1 else (float(v[0]) if v.dtype == torch.float64 else int(v[0]))
        return out


def get_train_val_split(data: SyntheticTwoViewDataset, **kwargs) -> Tuple[DataLoader, DataLoader]:
    """src/data_loader/utils.py:225-275 (single-dataset branch): the training loader shuffles, the validation
    loader serves a shallow copy of the data set switched to its validation indices."""
    data.is_training(True)
    val_data = copy.copy(data)
    val_data.is_training(False)
    return (DataLoader(data, **{**kwargs, "shuffle": True}), DataLoader(val_data, **{**kwargs, "shuffle": False}))
