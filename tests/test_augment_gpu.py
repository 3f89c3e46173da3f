"""GPU-side two-view augmentation (csrc/augment.cu, peclr_b200/gpu_augment.py) against golden
vectors produced by EXECUTING the reference's SampleAugmenter with OpenCV in the build container
(oracle/make_golden_augment.py -> tests/golden/augment.npz): rotate (cv2.warpAffine), crop, cv2.resize INTER_AREA
(down-scaling, up-scaling and clipped non-square crops), HSV colour jitter, ToTensor + Normalize.

Tolerances (8-bit images): the geometric stages (warp + crop + resize) are integer / fixed-order fp32 arithmetic and
must be BIT-EXACT; the colour jitter goes through OpenCV's fp32 HSV->BGR, where at most 0.1 % of the values may differ
by one level (operation order inside OpenCV's SIMD code); the normalised fp32 views follow within 1e-6 of
(level / 255 - mean) / std.
"""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
FLAG_NAMES = ("rotate", "crop", "random_crop", "resize", "color_jitter")
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "augment.npz"))


def _case(gold, name):
    from peclr_b200.gpu_augment import GpuTwoViewAugmenter

    flags = dict(zip(FLAG_NAMES, (bool(v) for v in gold[name + "_flags"])))
    rs = tuple(int(v) for v in gold[name + "_resize"])
    aug = GpuTwoViewAugmenter(flags, dict(resize_shape=rs))
    image, joints = gold[name + "_image"], gold[name + "_joints"]
    random.seed(int(gold[name + "_seed"]))
    drawn = aug.draw(joints[None], image.shape[:2])
    return aug, flags, rs, image, joints, drawn


def test_every_golden_case_matches_the_reference_augmenter(gold):
    from peclr_b200.gpu_augment import two_view_augment

    for name in (str(c) for c in gold["cases"]):
        aug, flags, rs, image, joints, drawn = _case(gold, name)
        h, w = image.shape[:2]
        src = torch.tensor(image).cuda()
        for jitter_on in ((False, True) if flags["color_jitter"] else (False,)):
            tab = aug.table(drawn, 1, h, w)
            if not jitter_on:
                tab["jitter"] = 0
            out = torch.empty((2, 3, rs[1], rs[0]), device="cuda")
            stage = torch.empty((2, rs[1], rs[0], 3), dtype=torch.uint8, device="cuda")
            two_view_augment(src, torch.from_numpy(tab.view(np.uint8)).cuda(), out, stage_out=stage)
            torch.cuda.synchronize()
            for v in (0, 1):
                want = gold[f"{name}_v{v + 1}_" + ("final" if jitter_on or not flags["color_jitter"] else "resized")]
                got = stage[v].cpu().numpy()
                diff = np.abs(got.astype(int) - want.astype(int))
                if jitter_on:
                    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3, (name, v, diff.max(), (diff != 0).mean())
                else:  # rotate + crop + resize: bit-exact
                    assert diff.max() == 0, (name, v, int(diff.max()), float((diff != 0).mean()))
                # ToTensor + Normalize of the 8-bit image the kernel produced
                t = torch.tensor(got).permute(2, 0, 1).float().div(255)
                t = (t - torch.tensor(MEAN)[:, None, None]) / torch.tensor(STD)[:, None, None]
                assert torch.allclose(out[v].cpu(), t, atol=1e-6, rtol=1e-6), name


def test_rotation_stage_alone_is_bit_exact(gold):
    """A full-frame 'crop' at the output size isolates cv2.warpAffine (fixed-point bilinear, zero border)."""
    from peclr_b200.gpu_augment import VIEW_DTYPE, invert_affine, rotation_matrix_2d, two_view_augment

    for name, v in (("fh224_down", 1), ("wide_320x240", 2)):
        aug, flags, rs, image, joints, drawn = _case(gold, name)
        d = drawn[0][v - 1]
        h, w = image.shape[:2]
        tab = np.zeros(1, dtype=VIEW_DTYPE)
        tab["m"][0] = invert_affine(d["m_fwd"])
        tab["sh"], tab["sw"], tab["cw"], tab["ch"], tab["rotate"] = h, w, w, h, 1
        out = torch.empty((1, 3, h, w), device="cuda")
        stage = torch.empty((1, h, w, 3), dtype=torch.uint8, device="cuda")
        two_view_augment(torch.tensor(image).cuda(), torch.from_numpy(tab.view(np.uint8)).cuda(), out, stage_out=stage)
        torch.cuda.synchronize()
        assert np.array_equal(stage[0].cpu().numpy(), gold[f"{name}_v{v}_rotated"]), name
    m = rotation_matrix_2d((100, 90), 17.0)  # cv2.getRotationMatrix2D values
    assert np.allclose(m, [[0.95630476, 0.2923717, -21.94392902], [-0.2923717, 0.95630476, 33.16970041]], atol=1e-7)


def test_augmenter_feeds_the_training_step(gold):
    """uint8 images + joints -> batch dict of the reference's schema -> one CUDA training step."""
    from oracle import peclr_oracle as po
    from peclr_b200.easydict import EasyDict
    from peclr_b200.gpu_augment import GpuTwoViewAugmenter
    from peclr_b200.hybrid2_model import Hybrid2Model

    name = "fh224_down"
    image, joints = gold[name + "_image"], gold[name + "_joints"]
    b = 4
    images = torch.tensor(np.stack([np.roll(image, 7 * i, axis=1) for i in range(b)])).pin_memory()
    jts = np.stack([joints + np.float32(i) for i in range(b)])
    flags = dict(rotate=True, crop=True, random_crop=True, resize=True, color_jitter=True)
    aug = GpuTwoViewAugmenter(flags, dict(resize_shape=(64, 64)))
    random.seed(11)
    batch = aug(images, jts)
    ref_keys = set(po.synthetic_batch(2, 64).keys())
    assert ref_keys <= set(batch)  # transformed_image1/2, angle_k, jitter_x_k, jitter_y_k (+ the colour factors)
    assert batch["transformed_image1"].shape == (b, 3, 64, 64) and batch["transformed_image1"].dtype == torch.float32
    assert batch["angle_1"].dtype == torch.float64 and batch["jitter_x_2"].dtype == torch.int64
    assert float(batch["angle_1"].abs().max()) <= 45 and int(batch["jitter_x_1"].max()) <= 0
    cfg = po.default_config(resnet_size="18", batch_size=b, num_samples=b * 8, projection_head_input_dim=512)
    model = Hybrid2Model(EasyDict(dict(cfg))).cuda()
    model.train()
    out = model.training_step(batch, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out["loss"]) and float(model.engine.grads.abs().max()) > 0
