"""Multi-GPU parity checks (one rank per GPU, NCCL).  Used three ways: `torchrun ... scripts/dist_check.py [B]`
stand-alone, from tests/test_multigpu.py, and by bench.py at world > 1 BEFORE its timed region, so that every
multi-GPU bench line carries its own parity verdict (`"parity": {...}`).

The oracle (oracle/peclr_oracle.py) is the checker here, never the thing measured:

* loss_kernel_parity: every rank owns B pairs of a seeded global batch; the fused kernel (peer stores over NVLink
  + flag barrier inside the launch) must reproduce, on every rank, the loss of the reference chain on the
  CONCATENATED global batch (z ordered [z1 of rank 0..R-1, z2 of rank 0..R-1]) and its gradient
  w.r.t. the rank's own rows -- eagerly and replayed from a CUDA graph (device-side launch counter / double-buffered
  z buffers).
* dp_step_parity: one whole data-parallel step (trunk + head + fused all-gather loss + backward + gradient
  all-reduce SUM) at B pairs per rank against SURVEY 8(e)'s single-process oracle: R independent oracle trunks /
  heads on the R shards with the same weights (BatchNorm statistics stay per shard, as in the reference, which has
  no SyncBN), z concatenated in the order above, vanila_contrastive_loss on the global batch, autograd back into
  every shard, gradients summed.  Tolerances: the fixed step tolerances of tests/parity_util.py.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOL_LOSS_REL, TOL_GRAD_REL = 1e-5, 1e-4       # fused loss kernel vs fp64 closed form
TOL_DLOSS, TOL_COS_ALL, TOL_COS_TOP = 1e-2, 0.93, 0.985  # bf16 trunk step vs fp32 oracle at B = 8 per rank (tests/parity_util.py)


def _max_over_ranks(x, dev):
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def loss_kernel_parity(engine, b, world, rank, dev, graph_replays=3):
    from oracle import peclr_oracle as po

    rng = np.random.RandomState(0)
    n = 2 * b * world
    p = rng.randn(n, 128).astype(np.float32)
    p[n // 2:] = p[: n // 2] + 0.5 * rng.randn(n // 2, 128).astype(np.float32)
    angle = np.floor(rng.uniform(-45, 45, n))
    jx, jy = -rng.randint(0, 15, n), -rng.randint(0, 15, n)
    ref = po.loss_chain_numpy(p, angle, jx, jy, (224, 224), True, True, dtype=np.float64)
    # local rows: view-1 rows [rank*b, (rank+1)*b) and view-2 rows [world*b + rank*b, ...)
    idx = np.concatenate([np.arange(rank * b, (rank + 1) * b), world * b + np.arange(rank * b, (rank + 1) * b)])
    tp = torch.tensor(p[idx], device=dev)
    ta, tx, ty = (torch.tensor(v[idx], device=dev) for v in (angle, jx, jy))
    gscale = np.abs(ref["g_p"]).max()
    worst_l = worst_g = 0.0

    def check(loss, g):
        nonlocal worst_l, worst_g
        worst_l = max(worst_l, abs(loss.item() - ref["loss"]) / abs(ref["loss"]))
        worst_g = max(worst_g, float(np.abs(g.cpu().numpy() - ref["g_p"][idx]).max() / gscale))

    first = None
    for _ in range(2):
        loss, stats, g = engine.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
        torch.cuda.synchronize()
        check(loss, g)
        first = first if first is not None else (loss.clone(), g.clone())
    # CUDA-graph replay of the same launch
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        engine.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(graph):
        loss, stats, g = engine.forward_loss(tp, ta, tx, ty, (224, 224), True, True)
    reproducible = True
    for _ in range(graph_replays):
        graph.replay()
        torch.cuda.synchronize()
        check(loss, g)
        reproducible &= bool(torch.equal(loss, first[0]) and torch.equal(g, first[1]))
    del graph
    worst_l, worst_g = _max_over_ranks(worst_l, dev), _max_over_ranks(worst_g, dev)
    repro = _max_over_ranks(0.0 if reproducible else 1.0, dev) == 0.0
    return {"dist_loss_rel": worst_l, "dist_grad_rel": worst_g, "global_rows": n, "bit_reproducible": repro,
            "pass": bool(worst_l <= TOL_LOSS_REL and worst_g <= TOL_GRAD_REL)}


def _group_of(name):
    if name.startswith("projection_head"):
        return "head"
    return {"0": "stem", "1": "stem", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}[name.split(".")[2]]


def dp_step_parity(world, rank, dev, b=8, size=64, warm_steps=300):
    from oracle import peclr_oracle as po
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    old_det = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        cfg = po.default_config(resnet_size="50", batch_size=b, num_samples=b * world * 64)
        torch.manual_seed(0)
        oracle = po.OracleHybrid2Model(cfg).to(dev)
        oracle.train()
        if rank == 0:  # warm start away from the chaotic default init, then share the weights
            opt = torch.optim.Adam(oracle.parameters(), lr=1e-3)
            for i in range(warm_steps):
                batch = {k: v.to(dev) for k, v in po.synthetic_batch(b, size, seed=100 + i).items()}
                opt.zero_grad(set_to_none=True)
                oracle.training_step(batch, i)["loss"].backward()
                opt.step()
            oracle.zero_grad(set_to_none=True)
        for t in oracle.state_dict().values():
            dist.broadcast(t, src=0)
        oracle.train_metrics, oracle.plot_params = {}, {}
        sd = {k: v.clone() for k, v in oracle.state_dict().items()}
        ours = Hybrid2Model(EasyDict(dict(cfg)))
        ours.load_state_dict({k: v.cpu() for k, v in sd.items()})
        ours.cuda()
        ours.engine.world, ours.engine.rank = world, rank
        shards = [{k: v.to(dev) for k, v in po.synthetic_batch(b, size, seed=500 + r).items()} for r in range(world)]
        ours.train()
        ours.zero_grad()
        out = ours.training_step(shards[rank], 0)
        out["loss"].backward()
        ours.sync_gradients()
        torch.cuda.synchronize()
        res = {"pass": True}
        # the same step with the gradient all-reduce overlapped with backward (one collective per ResNet stage on a
        # communication stream) must give the monolithic all-reduce's gradients
        mono = ours.engine.grads.clone()
        ours.zero_grad()
        ours.enable_overlapped_sync(True)
        ours.training_step(shards[rank], 0)["loss"].backward()
        ours.sync_gradients()
        ours.enable_overlapped_sync(False)
        torch.cuda.synchronize()
        ov = float((ours.engine.grads - mono).abs().max() / mono.abs().max())
        res["overlapped_allreduce_rel_diff"] = _max_over_ranks(ov, dev)
        res["pass"] = res["overlapped_allreduce_rel_diff"] <= 1e-6
        if rank == 0:
            # SURVEY 8(e) oracle: R shards through the same weights (BN statistics per shard), global-batch NT-Xent
            # (the running statistics move from shard to shard; the training-mode forward does not read them)
            z1s, z2s = [], []
            for r in range(world):
                z1, z2 = oracle.get_transformed_projections({k: v.clone() for k, v in shards[r].items()})
                z1s.append(z1), z2s.append(z2)
            loss = po.vanila_contrastive_loss(torch.cat(z1s), torch.cat(z2s))
            loss.backward()
            ref, got = {}, {}
            mine = dict(ours.named_parameters())
            for name, prm in oracle.named_parameters():
                if prm.grad is None:
                    continue
                ref.setdefault(_group_of(name), []).append(prm.grad.double().flatten())
                got.setdefault(_group_of(name), []).append(mine[name].grad.double().flatten())
            cosines = {}
            for k in ref:
                a, c = torch.cat(got[k]), torch.cat(ref[k])
                cosines[k] = float(a @ c / (a.norm() * c.norm()))
            a = torch.cat([torch.cat(v) for v in got.values()])
            c = torch.cat([torch.cat(v) for v in ref.values()])
            cosines["all"] = float(a @ c / (a.norm() * c.norm()))
            dloss = abs(float(out["loss"]) - float(loss))
            ok = dloss <= TOL_DLOSS and cosines["all"] >= TOL_COS_ALL and cosines["layer4"] >= TOL_COS_TOP and \
                cosines["head"] >= TOL_COS_TOP
            res.update({"dp_loss_ours": float(out["loss"]), "dp_loss_oracle": float(loss), "dp_dloss": dloss,
                        "dp_grad_cos": {k: round(v, 4) for k, v in cosines.items()},
                        "pass": bool(ok and res["pass"])})
        # the summed gradient must be the same on every rank after the all-reduce
        g = ours.engine.grads
        g0 = g.clone()
        dist.broadcast(g0, src=0)
        same = _max_over_ranks(float((g - g0).abs().max()), dev) == 0.0
        res["grads_identical_across_ranks"] = same
        flag = _max_over_ranks(0.0 if (res["pass"] and same) else 1.0, dev) == 0.0
        res["pass"] = bool(flag)
        del ours, oracle
        import gc

        gc.collect()  # (the model <-> engine cycle holds a symmetric-memory handle: free it here, not during a capture)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return res
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = old_det


def run_all(engine, b, world, rank, dev):
    """Both checks; returns the dict bench.py prints as "parity" (rank 0 holds the step-level details)."""
    out = loss_kernel_parity(engine, b, world, rank, dev)
    step = dp_step_parity(world, rank, dev)
    out.update({k: v for k, v in step.items() if k != "pass"})
    out["pass"] = bool(out["pass"] and step["pass"])
    out["tolerances"] = {"dist_loss_rel": TOL_LOSS_REL, "dist_grad_rel": TOL_GRAD_REL, "dp_dloss": TOL_DLOSS,
                         "dp_grad_cos_all": TOL_COS_ALL, "dp_grad_cos_layer4_head": TOL_COS_TOP}
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("PECLR_ALLOW_RANDOM_INIT", "1")
    from oracle import peclr_oracle as po
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cfg = po.default_config(resnet_size="18", batch_size=b, num_samples=b * 64)
    cfg["projection_head_input_dim"] = 512
    model = Hybrid2Model(EasyDict(dict(cfg))).cuda()
    model.engine.world, model.engine.rank = world, rank
    res = run_all(model.engine, b, world, rank, dev)
    if rank == 0:
        print(res)
        print("DIST_CHECK", "PASS" if res["pass"] else "FAIL", "world", world, "B", b)
    dist.destroy_process_group()
    sys.exit(0 if res["pass"] else 1)


if __name__ == "__main__":
    main()
