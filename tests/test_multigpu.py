"""N > 1 paths.  CPU (gloo, world 2): the global-batch ordering / gradient-sum semantics of the data-parallel
step against the single-process oracle.  GPU (needs >= 2 devices): the fused all-gather kernel under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, b, out):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import peclr_oracle as po
    from dist_rows import global_rows, local_rows

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    n = 2 * b * world
    p = rng.randn(n, 128)
    idx = local_rows(rank, world, b)
    # each rank runs the per-row part of the chain on its own rows, gathers z, computes the global loss
    loc = po.loss_chain_numpy(p[idx], None, None, None, (64, 64), False, False, dtype=np.float64)
    z_loc = torch.tensor(loc["z"])
    gathered = [torch.empty_like(z_loc) for _ in range(world)]
    dist.all_gather(gathered, z_loc)
    z_glob = np.zeros((n, 128))
    for r in range(world):
        z_glob[global_rows(r, world, b)] = gathered[r].numpy()
    glob = po.loss_chain_numpy(z_glob, None, None, None, (64, 64), False, False, dtype=np.float64)
    # gradient of the GLOBAL mean loss w.r.t. this rank's rows, summed over ranks == single-process gradient
    g = torch.zeros(n, 128, dtype=torch.float64)
    g[idx] = torch.tensor(glob["g_p"][idx])
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(out, loss=glob["loss"], g=g.numpy())
    dist.destroy_process_group()


def test_global_batch_semantics_gloo(tmp_path):
    import torch.multiprocessing as mp

    from oracle import peclr_oracle as po

    b, world = 4, 2
    out = str(tmp_path / "r.npz")
    mp.spawn(_gloo_worker, args=(world, 29613, b, out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.RandomState(0)
    p = rng.randn(2 * b * world, 128)
    z = po.loss_chain_numpy(p, None, None, None, (64, 64), False, False, dtype=np.float64)
    ref = po.loss_chain_numpy(z["z"], None, None, None, (64, 64), False, False, dtype=np.float64)
    assert abs(got["loss"] - ref["loss"]) < 1e-12
    assert np.abs(got["g"] - ref["g_p"]).max() < 1e-12


def test_row_mapping():
    from dist_rows import global_rows, local_rows

    world, b = 4, 3
    seen = np.concatenate([global_rows(r, world, b) for r in range(world)])
    assert sorted(seen) == list(range(2 * b * world))
    for r in range(world):
        g = global_rows(r, world, b)
        assert list(g[:b]) == list(range(r * b, (r + 1) * b))  # view-1 block of rank r
        assert list(g[b:]) == list(range(world * b + r * b, world * b + (r + 1) * b))  # its positives
        assert list(local_rows(r, world, b)) == list(g)


@pytest.mark.gpu
def test_fused_allgather_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29614", os.path.join(ROOT, "scripts", "dist_check.py"), "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
