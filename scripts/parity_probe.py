"""Exploration (GPU): step parity of the CUDA path vs the fp32 oracle on the GPU for several trunks / warm-start
lengths, next to what the reference arithmetic gives under torch bf16 autocast.  Prints a table; asserts nothing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PECLR_ALLOW_RANDOM_INIT", "1")

import parity_util as pu  # noqa: E402
from oracle import peclr_oracle as po  # noqa: E402


def run(size, b, hw, warm_b, steps, head_in=2048, aug=("crop", "rotate"), lr=1e-3):
    cfg = po.default_config(resnet_size=size, batch_size=b, num_samples=b * 64, projection_head_input_dim=head_in,
                            augmentation=aug)
    oracle = pu.warm_started_oracle(cfg, steps=steps, batch_size=warm_b, size=hw, lr=lr)
    cls = None
    if not aug:
        from peclr_b200.simclr_model import SimCLR as cls
    ours = pu.candidate_from(oracle, cfg, cls=cls)
    batch = pu.to_cuda(po.synthetic_batch(b, hw, seed=5))
    ref, ref_g = pu.oracle_step_on_gpu(oracle, batch)
    env = pu.autocast_envelope(oracle, batch, crop="crop" in aug, rotate="rotate" in aug)
    got, got_g = pu.candidate_step(ours, batch)
    if not aug:
        ref = {"loss": ref["loss"]}
    tag = "RN%s B=%d %d^2 warm %d@B%d lr %g aug=%s" % (size, b, hw, steps, warm_b, lr, ",".join(aug))
    try:
        pu.report_and_check(tag, got, got_g, ref, ref_g, check_stats=False, envelope=env, tol=(1.0, -1.0, -1.0))
    except AssertionError as e:
        print("  assertion:", e)
    del oracle, ours
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["a"]
    if "a" in which:
        run("50", 8, 64, 8, 100)
        run("50", 8, 64, 8, 300)
        run("50", 8, 64, 8, 150, aug=())
        run("50", 8, 64, 8, 300, aug=())
        run("101", 8, 64, 8, 150)
        run("101", 8, 64, 8, 400)
        run("152", 8, 64, 8, 200)
        run("152", 8, 64, 8, 500)
        run("152", 16, 224, 16, 100)
        run("152", 32, 224, 16, 100)
        run("152", 8, 64, 8, 300, lr=3e-4)
