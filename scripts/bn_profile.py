"""In-step (eager, warm L2) CUDA-event times of the BatchNorm kernels of one ResNet-50 step, per tensor shape -- the
ncu launch list times them with flushed caches, which overstates the small launches.  GPU box:
    python scripts/bn_profile.py [50|152] > gpurun_out/bn_profile.txt"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PECLR_ALLOW_RANDOM_INIT", "1")
import torch  # noqa: E402

import bench  # noqa: E402
from peclr_b200 import _lib  # noqa: E402
from peclr_b200.hybrid2_model import Hybrid2Model  # noqa: E402

NAMES = ("peclr_bn_apply", "peclr_bn_bwd_apply", "peclr_bn_bwd_reduce", "peclr_stem_bn_relu_pool",
         "peclr_stem_pool_bwd", "peclr_stem_input")


def main(model_size="50", batch=128, size=224):
    torch.cuda.set_device(0)
    from peclr_b200.synthetic import synthetic_batch

    cfg = bench.build_config(model_size, batch, 1, 1)
    model = Hybrid2Model(cfg).cuda()

    class _T:
        world_size, max_epochs = 1, 100

    model.trainer = _T()
    model.setup("fit")
    (opt,), _ = model.configure_optimizers()
    model.train()
    data = {k: v.cuda() for k, v in synthetic_batch(batch, size, seed=5, structured=False).items()}

    def step():
        out = model.training_step(data, 0)
        out["loss"].backward()

    for _ in range(3):
        step()
        opt.zero_grad()
    torch.cuda.synchronize()
    prof = _lib.profile_calls(step, NAMES)
    rows = collections.OrderedDict()
    for name, a, ms in prof:
        ints = tuple(int(x) for x in a if isinstance(x, int) and 0 < x < (1 << 31))[:3]
        r = rows.setdefault((name.replace("peclr_", ""), ints), [0, 0.0])
        r[0] += 1
        r[1] += ms
    tot = sum(v[1] for v in rows.values())
    print("%-22s %-28s %3s %9s %9s" % ("call", "ints (M, C, ...)", "n", "us/call", "total us"))
    for (name, ints), (n, ms) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        print("%-22s %-28s %3d %9.1f %9.0f" % (name, ",".join(map(str, ints)), n, ms / n * 1e3, ms * 1e3))
    print("total %.3f ms" % tot)


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["50"]))
