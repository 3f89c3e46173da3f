"""Multi-GPU parity of the fused embedding all-gather + NT-Xent (launch with torchrun, one rank per GPU).

Every rank owns B pairs of a seeded global batch; the fused kernel (peer stores over NVLink + flag barrier inside
the launch) must reproduce, on every rank, the loss of the reference chain on the CONCATENATED global batch
(z ordered [z1 of rank 0..R-1, z2 of rank 0..R-1]) and its gradient w.r.t. the rank's own rows.
Also replays the kernel from a CUDA graph several times (device-side launch counter / double-buffered z)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import peclr_oracle as po
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rng = np.random.RandomState(0)
    n = 2 * b * world
    p = rng.randn(n, 128).astype(np.float32)
    p[n // 2:] = p[: n // 2] + 0.5 * rng.randn(n // 2, 128).astype(np.float32)
    angle = np.floor(rng.uniform(-45, 45, n))
    jx, jy = -rng.randint(0, 15, n), -rng.randint(0, 15, n)
    ref = po.loss_chain_numpy(p, angle, jx, jy, (224, 224), True, True, dtype=np.float64)
    # local rows: view-1 rows [rank*b, (rank+1)*b) and view-2 rows [world*b + rank*b, ...)
    idx = np.concatenate([np.arange(rank * b, (rank + 1) * b), world * b + np.arange(rank * b, (rank + 1) * b)])
    cfg = po.default_config(resnet_size="18", batch_size=b, num_samples=b * 64)
    cfg["projection_head_input_dim"] = 512
    model = Hybrid2Model(EasyDict(dict(cfg))).cuda()
    eng = model.engine
    eng.world, eng.rank = world, rank
    dev = torch.device("cuda", local)
    tp = torch.tensor(p[idx], device=dev)
    ta, tx, ty = (torch.tensor(v[idx], device=dev) for v in (angle, jx, jy))
    ok = True
    for it in range(3):
        loss, stats, g = eng.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
        torch.cuda.synchronize()
        dl = abs(loss.item() - ref["loss"])
        dg = np.abs(g.cpu().numpy() - ref["g_p"][idx]).max() / np.abs(ref["g_p"]).max()
        ok &= dl <= 1e-5 * abs(ref["loss"]) and dg <= 1e-4
        if rank == 0:
            print("eager it %d: loss %.6f ref %.6f  grad rel err %.2e" % (it, loss.item(), ref["loss"], dg))
    # CUDA-graph replay of the same launch
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(graph):
        loss, stats, g = eng.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
    for it in range(4):
        graph.replay()
        torch.cuda.synchronize()
        dl = abs(loss.item() - ref["loss"])
        dg = np.abs(g.cpu().numpy() - ref["g_p"][idx]).max() / np.abs(ref["g_p"]).max()
        ok &= dl <= 1e-5 * abs(ref["loss"]) and dg <= 1e-4
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", "world", world, "B", b)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
