"""Device time of the two-view augmentation kernel (GPU only): B = 128 raw 224 x 224 images -> 2 x 128 views."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peclr_b200.gpu_augment import GpuTwoViewAugmenter, two_view_augment  # noqa: E402


def main():
    b = 128
    flags = dict(rotate=True, crop=True, random_crop=True, resize=True, color_jitter=True)
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (b, 224, 224, 3), dtype=torch.uint8, generator=g).cuda()
    for size, spread in ((224, 0.12), (128, 0.12), (224, 0.22)):
        rs = np.random.RandomState(1)
        joints = (rs.randn(b, 21, 3) * 224 * spread + 112).astype(np.float32)
        aug = GpuTwoViewAugmenter(flags, dict(resize_shape=(size, size)), rng=random.Random(3))
        from peclr_b200.gpu_augment import draw_batch_params

        bp = draw_batch_params(joints, (224, 224), flags, aug.params, aug.rng)
        tab = torch.from_numpy(aug.batch_table(bp, b, 224, 224).view(np.uint8)).cuda()
        out = torch.empty((2 * b, 3, size, size), device="cuda")
        for _ in range(3):
            two_view_augment(images, tab, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            two_view_augment(images, tab, out)
        e1.record()
        torch.cuda.synchronize()
        print("aug kernel: %d views -> %dx%d, mean crop side %.0f px: %.3f ms per batch" %
              (2 * b, size, size, float(bp["cw"].mean()), e0.elapsed_time(e1) / 20))


if __name__ == "__main__":
    main()
