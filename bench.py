#!/usr/bin/env python
"""Benchmark of the PeCLR pre-training step (BASELINE.json: two-view images/sec, ResNet-50, batch 128, 224^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model 50|152] [--batch B]

ours       one process per GPU (torchrun for N > 1); a step = Hybrid2Model.training_step + backward + gradient
           exchange + fused LARS-Adam step on one synthetic two-view batch.  Prints ONE JSON line (rank 0).
reference  the reference's own CPU path (oracle port of Hybrid2Model.training_step / backward / LARS-Adam step;
           the reference is Python and cannot travel to the GPU box) on all host cores, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# conv-only training FLOPs per image, 2*(3*MAC_fwd - MAC_conv1)
F_TRAIN = {("50", 224): 24.287e9, ("152", 224): 68.834e9, ("50", 64): 1.983e9, ("50", 128): 7.930e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def ncu_traffic(args):
    """dram__bytes_read + dram__bytes_write per conv_gemm_kernel launch (average over the launches of one step),
    from the committed ncu capture (profiles/roofline_r01.json, taken on the default ResNet-50 workload); None for
    any other workload or if the capture is missing."""
    if (args.model, args.batch, args.size) != ("50", 128, 224):
        return None
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "roofline_r01.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            time.sleep(0.03)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def build_config(args, world):
    from peclr_b200.easydict import EasyDict

    return EasyDict(batch_size=args.batch, lr=1e-4, opt_weight_decay=1e-6, output_dim=128,
                    projection_head_hidden_dim=512, projection_head_input_dim=2048, warmup_epochs=10,
                    num_of_mini_batch=1, augmentation=["crop", "rotate"], optimizer="LARS",
                    resnet_size=args.model, num_samples=args.batch * world * 1000)


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from peclr_b200 import _lib
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.lightning import seed_everything
    from peclr_b200.synthetic import synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seed_everything(5)
    cfg = build_config(args, world)
    model = Hybrid2Model(cfg)
    model.cuda()

    class _T:
        world_size, max_epochs = world, 100

    model.trainer = _T()
    model.engine.world, model.engine.rank = world, rank
    model.setup("fit")
    (opt,), (sch,) = model.configure_optimizers()
    sched = sch["scheduler"]
    model.train()

    # two distinct host batches (pinned) so consecutive steps never see the same input
    host = [synthetic_batch(args.batch, args.size, seed=5 + 17 * rank + i, structured=False, pin_memory=True)
            for i in range(2)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]

    def eager_step(batch):
        opt.zero_grad()
        out = model.training_step(batch, 0)
        out["loss"].backward()
        model.sync_gradients()
        opt.step()
        sched.step()
        return out["loss"]

    graphed = None
    if not args.no_graph:
        from peclr_b200.graphed import GraphedStep

        graphed = GraphedStep(model, resident[0], grad_scale=1.0)

    def step(batch):
        """One optimiser step.  With the CUDA graph, training_step + backward are replayed from the capture
        (same kernels, one submission); gradient exchange and the fused optimiser run after it."""
        if graphed is None:
            return eager_step(batch)
        opt.zero_grad()
        out = graphed(batch)
        model.sync_gradients()
        opt.step()
        sched.step()
        return out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    for i in range(args.warmup):
        step(resident[i % 2])
    barrier()
    if args.ncu_step:
        # launch-list capture: `ncu --profile-from-start off ... bench.py --ncu-step --no-graph` sees exactly the
        # launches of ONE step (cudaProfilerStart/Stop); nothing printed under a profiler is a bench value
        torch.cuda.profiler.start()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    launches0 = _lib.LAUNCHES
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = step(resident[i % 2])
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = _lib.LAUNCHES - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    images = 2 * args.batch * world * args.steps
    value = images / (ms / 1e3)
    final_loss = float(loss.item())

    # ---- end to end: host (pinned) batch -> H2D on a copy stream (double buffered) -> step -> loss D2H ----
    copy_stream = torch.cuda.Stream()
    staged = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k, v in host[i % 2].items():
                staged[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(n):
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        stage(0)
        for i in range(n):
            if i + 1 < n:
                stage(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            l = step(staged[i % 2])
            consumed[i % 2].record(torch.cuda.current_stream())
            loss_host.copy_(l.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_run(max(3, args.warmup))
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    e2e_run(args.steps)
    ev1.record()
    barrier()
    # the host is part of this path: take the larger of the device-event time and the wall clock
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = images / (float(t.item()) / 1e3)

    # ---- dominant-kernel roofline: every tensor-core conv launch of ONE step bracketed by CUDA events ----
    roof = None
    cpu = None
    # (a step is collective when world > 1 -- fused all-gather, gradient all-reduce -- so every rank runs it)
    prof = _lib.profile_calls(lambda: eager_step(resident[0]),
                              {"peclr_conv2d_fprop", "peclr_conv2d_dgrad", "peclr_conv2d_dgrad_bnreduce",
                               "peclr_conv2d_wgrad", "peclr_stem_fprop", "peclr_stem_wgrad"})
    if rank == 0:
        pk = peaks()
        flops = {"gemm": 0.0, "wgrad": 0.0}
        times = {"gemm": 0.0, "wgrad": 0.0}
        for name, a, ms_k in prof:
            kind = "wgrad" if "wgrad" in name else "gemm"
            flops[kind] += conv_flops(name, a)
            times[kind] += ms_k
        ach = flops["gemm"] / (times["gemm"] * 1e-3) / 1e12 if times["gemm"] > 0 else 0.0
        ach_w = flops["wgrad"] / (times["wgrad"] * 1e-3) / 1e12 if times["wgrad"] > 0 else 0.0
        f_train = F_TRAIN.get((args.model, args.size))
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM fprop+dgrad)",
                "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4), "traffic": ncu_traffic(args),
                "peak_source": pk["src"] + " (sustained)",
                "kernel_ms_per_step": round(times["gemm"], 3),
                "wgrad_kernel": {"achieved": round(ach_w, 1), "frac": round(ach_w / pk["tf_sustained"], 4),
                                 "kernel_ms_per_step": round(times["wgrad"], 3)},
                "step_conv_flop_frac": round(value * f_train / world / (pk["tf_sustained"] * 1e12), 4) if f_train else None}
        try:  # (explanatory extra: never let it take the bench line down)
            roof["per_launch_roofline"] = launch_bound_fraction(prof, pk["tf_sustained"], pk["hbm"])
        except Exception as exc:  # pragma: no cover
            roof["per_launch_roofline"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args, budget_s=20.0)
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": "two-view images/sec, PeCLR pre-training step (fwd+bwd+LARS-Adam), ResNet-%s bs%d %dx%d"
                      % (args.model, args.batch, args.size, args.size),
            "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ResNet-%s PeCLR step, per-GPU batch %d (2x%d images), %dx%d synthetic two-view, "
                                   "crop+rotate equivariance, NT-Xent over the global batch, LARS-Adam"
                                   % (args.model, args.batch, args.batch, args.size, args.size),
                       "global_batch": args.batch * world, "parallelism": "dp%d" % world,
                       "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2; no explicit flush"
                             % (h2d_bytes / 1e6)},
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "final_loss": final_loss,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def conv_bytes(name, a):
    """Algorithmic HBM bytes of one conv launch from its C-ABI arguments: every operand once (bf16 activations and
    weights; a read-modify-write of dx when the dgrad accumulates; the BatchNorm input tile of the fused reduction)."""
    if name.startswith("peclr_stem"):
        n, h, w = a[3], a[4], a[5]
        return 2.0 * (n * (h // 2 + 3) * (w // 2 + 4) * 16 + n * (h // 2) * (w // 2) * 64 + 64 * 4 * 64)
    n, h, w, cin, cout, k, s = a[3:10]
    big, small = n * h * w * cin, n * (h // s) * (w // s) * cout  # input-side / output-side activation elements
    total = big + small + cout * k * k * cin
    if name == "peclr_conv2d_dgrad" and a[10]:
        total += big
    if name == "peclr_conv2d_dgrad_bnreduce":
        total += big
    return 2.0 * total


def launch_bound_fraction(prof, peak_tflops, peak_gbs):
    """Sum over the fprop / dgrad launches of max(FLOPs / tensor peak, algorithmic bytes / HBM peak), divided by the
    sum of their measured durations: how close the kernel runs to the roofline of each individual launch (many of
    them are 1x1 convolutions with K <= 256, which no kernel can run at the tensor peak)."""
    bound = meas = 0.0
    n_hbm = n_all = 0
    for name, a, ms_k in prof:
        if "wgrad" in name:
            continue
        t_tensor = conv_flops(name, a) / (peak_tflops * 1e12) * 1e3
        t_hbm = conv_bytes(name, a) / (peak_gbs * 1e9) * 1e3
        bound += max(t_tensor, t_hbm)
        meas += ms_k
        n_all += 1
        n_hbm += t_hbm > t_tensor
    if meas <= 0:
        return None
    return {"bound_ms": round(bound, 3), "measured_ms": round(meas, 3), "frac": round(bound / meas, 4),
            "launches": n_all, "hbm_bound_launches": n_hbm}


def conv_flops(name, a):
    """2*M*N*K of one conv launch from its C-ABI arguments."""
    if name.startswith("peclr_stem"):
        n, h, w = a[3], a[4], a[5]
        return 2.0 * n * (h // 2) * (w // 2) * 64 * 147
    n, h, w, cin, cout, k, s = a[3:10]
    return 2.0 * n * (h // s) * (w // s) * cout * cin * k * k


# ---------------------------------------------------------------------------------------------- CPU arms
def _oracle_model(args, batch):
    import torch

    from oracle import peclr_oracle as po

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = po.default_config(resnet_size=args.model, batch_size=batch, num_samples=batch * 1000)
    torch.manual_seed(5)
    model = po.OracleHybrid2Model(cfg)
    model.trainer = po._TrainerStub(world_size=1, max_epochs=100)
    model.setup("fit")
    (opt,), (sch,) = model.configure_optimizers()
    return po, model, opt, sch["scheduler"]


def _time_oracle_steps(args, batch, warm, steps):
    po, model, opt, sched = _oracle_model(args, batch)
    data = po.synthetic_batch(batch, args.size, seed=5, structured=False)
    for _ in range(warm):
        po.oracle_step(model, data, opt, sched)
    t0 = time.perf_counter()
    for _ in range(steps):
        po.oracle_step(model, data, opt, sched)
    return (time.perf_counter() - t0) / steps


def pick_sample_batch(args, target_s=2.5):
    """Largest per-step sample (pairs) whose step takes about target_s on this host."""
    t4 = _time_oracle_steps(args, 4, 1, 1)
    b = 4
    while b * 2 <= args.batch and t4 * (b * 2 / 4) <= target_s:
        b *= 2
    return b


def cpu_baseline(args, budget_s=20.0):
    b = pick_sample_batch(args, target_s=budget_s / 4)
    dt = _time_oracle_steps(args, b, 1, 2)
    return {"value": round(2 * b / dt, 2), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle port of the reference step (fp32, torch CPU, %d threads): 1 warm-up + 2 timed steps at "
                      "B=%d of %d pairs, %dx%d, scaled per image" % (os.cpu_count(), b, args.batch, args.size, args.size)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample per step so that the whole --steps K --warmup W run stays within ~2 minutes of CPU work
    b = pick_sample_batch(args, target_s=min(2.5, 120.0 / max(1, args.steps + args.warmup)))
    po, model, opt, sched = _oracle_model(args, b)
    data = po.synthetic_batch(b, args.size, seed=5, structured=False)
    for _ in range(args.warmup):
        po.oracle_step(model, data, opt, sched)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.oracle_step(model, data, opt, sched)
    dt = (time.perf_counter() - t0) / args.steps
    value = round(2 * b / dt, 2)
    sample = ("oracle port of the reference's Hybrid2Model.training_step + backward + LARSWrapper(Adam).step (fp32, "
              "torch CPU, %d threads); each step = B=%d pairs of the %d-pair batch, %dx%d, throughput per image"
              % (os.cpu_count(), b, args.batch, args.size, args.size))
    print(json.dumps({
        "impl": "reference",
        "metric": "two-view images/sec, PeCLR pre-training step (fwd+bwd+LARS-Adam), ResNet-%s bs%d %dx%d"
                  % (args.model, args.batch, args.size, args.size),
        "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-%s PeCLR step on host cores, bounded sample B=%d pairs per step" % (args.model, b)},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="50", choices=["18", "34", "50", "101", "152"])
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--ncu-step", dest="ncu_step", action="store_true",
                    help="run the warm-up, then ONE step between cudaProfilerStart/Stop, and exit (for ncu)")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="submit kernels eagerly (no CUDA graph)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
