#!/usr/bin/env python
"""Benchmark of the PeCLR pre-training step (BASELINE.json: two-view images/sec, ResNet-50, batch 128, 224^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model 50|152] [--batch B]
                    [--accumulate A]

ours       one process per GPU (torchrun for N > 1); a step = one OPTIMISER step: A micro-steps of
           Hybrid2Model.training_step + backward (A = --accumulate, Lightning's accumulate_grad_batches; default 1),
           then gradient exchange + fused LARS-Adam, each micro-step on its own synthetic two-view batch.  Prints ONE
           JSON line (rank 0).  On the default workload (BASELINE config 2: ResNet-50, B = 128, 224^2) the line also
           carries `secondary`: ResNet-152 B = 128 (config 4) and ResNet-152 B = 64 x accumulate 16 (config 5, the
           paper's recipe), measured the same way in the same process, and for N > 1 `parity`: the multi-GPU parity
           checks of scripts/dist_check.py run before the timed region (a failing check exits non-zero).
reference  the reference's own CPU path on all host cores, bounded sample per step: the reference's Hybrid2Model
           imported from /root/reference where that exists (build container; kind "reference"), else its oracle port
           (oracle/peclr_oracle.py, bit-identical to the executed reference -- tests/test_oracle_vs_reference.py; the
           reference is a Python tree that cannot travel to the GPU box; kind "port").
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
CONV_CALLS = {"peclr_conv2d_fprop", "peclr_conv2d_dgrad", "peclr_conv2d_dgrad_bnreduce", "peclr_conv2d_dgrad_finish",
              "peclr_conv2d_dgrad_finish_lattice",
              "peclr_conv2d_wgrad", "peclr_conv2d_wgrad_partials", "peclr_stem_fprop", "peclr_stem_wgrad"}
CPU_SAMPLE_PAIRS = 32  # per-step sample of BOTH CPU legs (cpu_baseline and --impl reference), scaled per image


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def lib_digest():
    """Digest of the CUDA sources + flags the loaded library was built from (peclr_b200/build.py stamp)."""
    try:
        return open(os.path.join(ROOT, "peclr_b200", "csrc", ".build_stamp")).read().strip()[:16]
    except OSError:
        return None


def ncu_traffic(model, batch, size):
    """dram__bytes_read + dram__bytes_write per conv_gemm_kernel launch (average over the launches of one step) from
    the committed ncu launch list -- used ONLY if that capture was taken on the build now loaded (same source
    digest) and on this workload; otherwise null (a stale constant would be worse than none)."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_r02.json")))
        if rec.get("lib_digest") == lib_digest() and rec.get("workload") == [model, batch, size]:
            return int(rec["dram_bytes_per_launch"])
    except Exception:
        pass
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


def build_config(model, batch, accumulate, world):
    from peclr_b200.easydict import EasyDict

    return EasyDict(batch_size=batch, lr=1e-4, opt_weight_decay=1e-6, output_dim=128,
                    projection_head_hidden_dim=512, projection_head_input_dim=2048, warmup_epochs=10,
                    num_of_mini_batch=accumulate, augmentation=["crop", "rotate"], optimizer="LARS",
                    resnet_size=model, num_samples=batch * world * accumulate * 1000)


def workload_name(model, batch, size, accumulate):
    acc = ", accumulate_grad_batches %d (one optimiser step per %d micro-batches)" % (accumulate, accumulate) \
        if accumulate > 1 else ""
    return ("ResNet-%s PeCLR step, per-GPU batch %d (2x%d images), %dx%d synthetic two-view, crop+rotate "
            "equivariance, NT-Xent over the global batch, LARS-Adam%s" % (model, batch, batch, size, size, acc))


# ---------------------------------------------------------------------------------------------- our arm
def measure(ctx, model_size, batch, size, accumulate, steps, warmup, sample_clocks=False, profile=True,
            ncu_step=False, no_graph=False):
    """Builds the model, times `steps` optimiser steps device-resident and end to end, profiles the conv launches of
    one micro-step.  Returns a dict (rank 0: complete; other ranks: partial)."""
    import torch
    import torch.distributed as dist

    from peclr_b200 import _lib
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.lightning import seed_everything
    from peclr_b200.synthetic import synthetic_batch

    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    seed_everything(5)
    cfg = build_config(model_size, batch, accumulate, world)
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

    # distinct host batches (pinned): consecutive micro-steps never see the same input
    n_host = 2 if accumulate == 1 else min(accumulate, 4)
    host = [synthetic_batch(batch, size, seed=5 + 17 * rank + i, structured=False, pin_memory=True)
            for i in range(n_host)]
    micro_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    scale = 1.0 / accumulate

    def eager_micro(mb):
        out = model.training_step(mb, 0)
        (out["loss"] * scale).backward() if accumulate > 1 else out["loss"].backward()
        return out["loss"]

    graphed = None
    if not no_graph:
        from peclr_b200.graphed import GraphedStep

        graphed = GraphedStep(model, resident[0], grad_scale=scale)

    def step(get_batch, after_micro=None):
        """One optimiser step = `accumulate` micro-steps (training_step + backward; with the CUDA graph replayed
        from the capture: same kernels, one submission each), then gradient exchange + the fused optimiser."""
        opt.zero_grad()
        synced = False
        for j in range(accumulate):
            mb = get_batch(j)
            closing = j == accumulate - 1  # (data parallel: this micro-step all-reduces its gradients as it goes)
            model.enable_overlapped_sync(closing)
            if graphed is not None:
                loss = graphed(mb, sync=closing)["loss"]
                synced = graphed.synced
            else:
                loss = eager_micro(mb)
            if after_micro is not None:
                after_micro(j, loss)
        if not synced:
            model.sync_gradients()
        opt.step()
        sched.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------------
    ctr = [0]

    def next_resident(_):
        ctr[0] += 1
        return resident[ctr[0] % n_host]

    for _ in range(warmup):
        step(next_resident)
    barrier()
    if ncu_step:
        # launch-list capture: `ncu --profile-from-start off ... bench.py --ncu-step --no-graph` sees exactly the
        # launches of ONE step (cudaProfilerStart/Stop); nothing printed under a profiler is a bench value
        torch.cuda.profiler.start()
        step(next_resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    launches0 = _lib.LAUNCHES
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(ctx["local"])
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        loss = step(next_resident)
    ev1.record()
    barrier()
    if sampler is not None:
        sampler.stop_flag = True
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = _lib.LAUNCHES - launches0
    images = 2 * batch * accumulate * world * steps
    value = images / (ms / 1e3)
    final_loss = float(loss.item())

    # ---- end to end ------------------------------------------------------------------------------------------
    e2e = None
    if ctx.get("e2e_input", "fp32") == "u8":
        e2e = e2e_from_raw_images(ctx, model, opt, sched, batch, size, accumulate, steps, images, barrier,
                                  max_over_ranks, no_graph)
    if e2e is None:
        # (fp32 path) host (pinned) two-view batch -> H2D on a copy stream (double buffered) -> step -> loss D2H
        copy_stream = torch.cuda.Stream()
        staged = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        def stage(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                for k, v in host[i % n_host].items():
                    staged[slot][k].copy_(v, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_run(n):
            total = n * accumulate
            cur = torch.cuda.current_stream()
            for s in range(2):
                consumed[s].record(cur)
            stage(0)
            it = [0]

            def get(_):
                i = it[0]
                if i + 1 < total:
                    stage(i + 1)
                cur.wait_event(ready[i % 2])
                return staged[i % 2]

            def after(_, l):
                i = it[0]
                consumed[i % 2].record(cur)
                loss_host.copy_(l.detach(), non_blocking=True)
                it[0] = i + 1

            for _ in range(n):
                step(get, after)
            cur.synchronize()

        e2e_run(3 if accumulate == 1 else 1)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        e2e_run(steps)
        ev1.record()
        barrier()
        # the host is part of this path: take the larger of the device-event time and the wall clock
        e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
        e2e = {"value": round(images / (e2e_ms / 1e3), 1), "unit": "images/s",
               "h2d_bytes_per_step": micro_bytes * accumulate, "d2h_bytes_per_step": 4 * accumulate,
               "input": "two normalised fp32 views per sample from pinned host memory"}

    # ---- dominant-kernel roofline: every tensor-core conv launch of ONE micro-step bracketed by CUDA events ----
    roof = None
    if profile:
        # (a step is collective when world > 1 -- fused all-gather -- so every rank runs it)
        opt.zero_grad()
        prof = _lib.profile_calls(lambda: eager_micro(resident[0]), CONV_CALLS)
        opt.zero_grad()
        if rank == 0:
            pk = peaks()
            flops = {"gemm": 0.0, "wgrad": 0.0}
            times = {"gemm": 0.0, "wgrad": 0.0}
            nbytes = 0.0       # algorithmic HBM bytes of the fprop / dgrad launches
            t_hbm_bound = 0.0  # time spent in launches whose own roofline is the HBM one
            for name, a, ms_k in prof:
                kind = "wgrad" if "wgrad" in name else "gemm"
                flops[kind] += conv_flops(name, a)
                times[kind] += ms_k
                if kind == "gemm":
                    nbytes += conv_bytes(name, a)
                    if conv_bytes(name, a) / (pk["hbm"] * 1e9) > conv_flops(name, a) / (pk["tf_sustained"] * 1e12):
                        t_hbm_bound += ms_k
            ach = flops["gemm"] / (times["gemm"] * 1e-3) / 1e12 if times["gemm"] > 0 else 0.0
            ach_b = nbytes / (times["gemm"] * 1e-3) / 1e9 if times["gemm"] > 0 else 0.0
            ach_w = flops["wgrad"] / (times["wgrad"] * 1e-3) / 1e12 if times["wgrad"] > 0 else 0.0
            f_train = F_TRAIN.get((model_size, size))
            hbm_share = t_hbm_bound / times["gemm"] if times["gemm"] > 0 else 0.0
            # the kernel class is a mix of HBM-bound launches (1x1 convolutions with K <= 256, the finishing dgrads) and
            # tensor-bound ones: the headline bound is the one most of its TIME is spent under; the other view and the
            # per-launch figure stand next to it
            tensor_view = {"achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                           "frac": round(ach / pk["tf_sustained"], 4)}
            hbm_view = {"achieved": round(ach_b, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(ach_b / pk["hbm"], 4)}
            head = dict(hbm_view, bound="hbm") if hbm_share >= 0.5 else dict(tensor_view, bound="tensor")
            roof = {"bound": head["bound"], "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM fprop+dgrad)",
                    "achieved": head["achieved"], "peak": head["peak"], "unit": head["unit"], "frac": head["frac"],
                    "traffic": ncu_traffic(model_size, batch, size),
                    "algorithmic_bytes_per_launch": round(nbytes / max(1, sum(1 for n_, _, _ in prof if "wgrad" not in n_))),
                    "hbm_bound_time_share": round(hbm_share, 3), "tensor_view": tensor_view, "hbm_view": hbm_view,
                    "peak_source": pk["src"] + " (sustained bf16; HBM copy bandwidth)",
                    "kernel_ms_per_micro_step": round(times["gemm"], 3),
                    "wgrad_kernel": {"achieved": round(ach_w, 1), "frac": round(ach_w / pk["tf_sustained"], 4),
                                     "kernel_ms_per_micro_step": round(times["wgrad"], 3),
                                     "note": "incl. the ordered reduction of the pixel splits"},
                    "step_conv_flop_frac": round(value * f_train / world / (pk["tf_sustained"] * 1e12), 4)
                    if f_train else None}
            try:  # (explanatory extra: never let it take the bench line down)
                roof["per_launch_roofline"] = launch_bound_fraction(prof, pk["tf_sustained"], pk["hbm"])
                if os.environ.get("PECLR_BENCH_DUMP_LAUNCHES"):
                    dump_launch_table(prof, pk["tf_sustained"], pk["hbm"], os.environ["PECLR_BENCH_DUMP_LAUNCHES"])
            except Exception as exc:  # pragma: no cover
                roof["per_launch_roofline"] = {"error": repr(exc)}
    res = {"value": round(value, 1), "ms_per_step": round(ms / steps, 3), "steps": steps, "warmup": warmup,
           "e2e": e2e,
           "gpu_launches": int(launches), "roofline": roof, "final_loss": final_loss,
           "clocks": sampler.summary() if sampler is not None else None, "micro_bytes": micro_bytes,
           "kernels_per_micro_step": graphed.kernels_per_replay if graphed is not None else None}
    del graphed, model, opt, sched, resident, host
    _release()
    return res


def e2e_from_raw_images(ctx, model, opt, sched, batch, size, accumulate, steps, images, barrier, max_over_ranks,
                        no_graph):
    """End to end through the public API a training loop uses on raw data: pinned uint8 images + joints on the host
    -> GpuTwoViewAugmenter (parameters drawn on the host, ONE upload of 8-bit images + parameter table on a copy
    stream, rotate / crop / resize / colour-jitter / normalise in one kernel) -> training_step + backward + optimiser
    -> loss read back.  The augmentation is EXTRA work inside the timed region compared with `value`; what it buys is
    the input format: one 8-bit source image per sample instead of two normalised fp32 views (8x fewer bytes over
    PCIe)."""
    import random

    import numpy as np
    import torch

    from peclr_b200.gpu_augment import GpuTwoViewAugmenter

    dev, rank = ctx["dev"], ctx["rank"]
    flags = dict(rotate=True, crop=True, random_crop=True, resize=True, color_jitter=True)  # README.md:51 recipe
    aug = GpuTwoViewAugmenter(flags, dict(resize_shape=(size, size)), device=dev, rng=random.Random(5 + rank))
    g = torch.Generator().manual_seed(50 + rank)
    n_host = 2 if accumulate == 1 else 4
    raw = [torch.randint(0, 256, (batch, size, size, 3), dtype=torch.uint8, generator=g).pin_memory()
           for _ in range(n_host)]
    rs = np.random.RandomState(60 + rank)
    joints = [(rs.randn(batch, 21, 3) * size * 0.12 + size * 0.5).astype(np.float32) for _ in range(n_host)]
    example = aug(raw[0], joints[0])
    torch.cuda.synchronize()
    scale = 1.0 / accumulate
    graphed = None
    if not no_graph:
        from peclr_b200.graphed import GraphedStep

        graphed = GraphedStep(model, example, grad_scale=scale)
    copy_stream = torch.cuda.Stream()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    small_bytes = aug._staging["slots"][0]["host"].numel()

    def prepare(i):
        """Upload + augmentation of micro-batch i on the copy stream (overlaps the compute of micro-batch i - 1)."""
        mb = aug(raw[i % n_host], joints[i % n_host], copy_stream=copy_stream, kernel_on_copy_stream=True)
        return mb, aug.last_ready

    def run(n):
        total = n * accumulate
        cur = torch.cuda.current_stream()
        nxt = prepare(0)
        i = 0
        for _ in range(n):
            opt.zero_grad()
            synced = False
            for j in range(accumulate):
                mb, ready = nxt
                if i + 1 < total:
                    nxt = prepare(i + 1)
                cur.wait_event(ready)
                closing = j == accumulate - 1
                model.enable_overlapped_sync(closing)
                if graphed is not None:
                    loss = graphed(mb, sync=closing)["loss"]
                    synced = graphed.synced
                else:
                    out = model.training_step(mb, 0)
                    (out["loss"] * scale).backward()
                    loss = out["loss"]
                loss_host.copy_(loss.detach(), non_blocking=True)
                i += 1
            if not synced:
                model.sync_gradients()
            opt.step()
            sched.step()
        cur.synchronize()

    run(3 if accumulate == 1 else 1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    run(steps)
    ev1.record()
    barrier()
    ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
    del graphed
    return {"value": round(images / (ms / 1e3), 1), "unit": "images/s",
            "h2d_bytes_per_step": (raw[0].numel() + small_bytes) * accumulate, "d2h_bytes_per_step": 4 * accumulate,
            "input": "raw uint8 images (one per sample) + 21 joints from pinned host memory; the two views are made on "
                     "the GPU (rotate, crop, INTER_AREA resize, HSV jitter, normalise) inside the timed region"}


def _release():
    import gc

    import torch

    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as dist

    os.environ.setdefault("PECLR_ALLOW_RANDOM_INIT", "1")  # synthetic benchmark: random init is the stated workload
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
    ctx = dict(world=world, rank=rank, local=local, dev=dev, e2e_input=args.e2e_input)

    # ---- multi-GPU parity BEFORE anything is timed (scripts/dist_check.py; the oracle is the checker) ----------
    parity = None
    if world > 1 and not args.no_parity and not args.ncu_step:
        from peclr_b200.hybrid2_model import Hybrid2Model
        from scripts import dist_check

        pcfg = build_config("18", args.batch, 1, world)  # (a small trunk: only its engine's loss path is used)
        pcfg.projection_head_input_dim = 512
        probe = Hybrid2Model(pcfg).cuda()
        probe.engine.world, probe.engine.rank = world, rank
        parity = dist_check.run_all(probe.engine, args.batch, world, rank, dev)
        del probe
        _release()  # (symmetric-memory handles must be gone before a CUDA graph is captured: freeing them is not
        # a capturable operation, and the model <-> engine reference cycle leaves that to the cyclic collector)
        if not parity["pass"]:
            if rank == 0:
                print(json.dumps({"parity": parity, "error": "multi-GPU parity check failed"}))
            dist.destroy_process_group()
            raise SystemExit(3)

    main_res = measure(ctx, args.model, args.batch, args.size, args.accumulate, args.steps, args.warmup,
                       sample_clocks=True, ncu_step=args.ncu_step, no_graph=args.no_graph)
    if args.ncu_step:
        return
    default_workload = (args.model, args.batch, args.size, args.accumulate) == ("50", 128, 224, 1)
    secondary = None
    if default_workload and not args.no_secondary:
        secondary = {}
        s152 = max(5, args.steps // 5)
        r = measure(ctx, "152", 128, 224, 1, s152, 3)
        secondary["rn152_bs128"] = dict(config={"workload": workload_name("152", 128, 224, 1),
                                                "baseline_config": "BASELINE.json configs[3] (per GPU)"}, **_slim(r))
        s5 = max(2, args.steps // 12)
        r = measure(ctx, "152", 64, 224, 16, s5, 1)
        secondary["rn152_bs64_acc16"] = dict(config={"workload": workload_name("152", 64, 224, 16),
                                                     "baseline_config": "BASELINE.json configs[4] (per GPU)"},
                                             **_slim(r))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, budget_s=20.0)
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": "two-view images/sec, PeCLR pre-training step (fwd+bwd+LARS-Adam), ResNet-%s bs%d %dx%d"
                      % (args.model, args.batch, args.size, args.size),
            "value": main_res["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.model, args.batch, args.size, args.accumulate),
                       "global_batch": args.batch * world, "parallelism": "dp%d" % world,
                       "l2": "inputs (%.0f MB/micro-step) and activations exceed the 126 MB L2; no explicit flush"
                             % (main_res["micro_bytes"] / 1e6),
                       "init": "random (torchvision default init; no network for ImageNet weights)"},
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"],
            "clocks": main_res["clocks"],
            "roofline": main_res["roofline"],
            "cpu_baseline": cpu,
            "final_loss": main_res["final_loss"],
            "lib_digest": lib_digest(),
        }
        if secondary is not None:
            line["secondary"] = secondary
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _slim(r):
    roof = r["roofline"] or {}
    return {"value": r["value"], "unit": "images/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
            "warmup": r["warmup"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "final_loss": r["final_loss"],
            "roofline": {k: roof.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac",
                                                  "step_conv_flop_frac", "wgrad_kernel")}}


def conv_bytes(name, a):
    """Algorithmic HBM bytes of one conv launch from its C-ABI arguments: every operand once (bf16 activations and
    weights; a read-modify-write of dx when the dgrad accumulates; the BatchNorm input tile of the fused reduction)."""
    if name.startswith("peclr_stem"):
        n, h, w = a[3], a[4], a[5]
        return 2.0 * (n * (h // 2 + 3) * (w // 2 + 4) * 16 + n * (h // 2) * (w // 2) * 64 + 64 * 4 * 64)
    if name.startswith("peclr_conv2d_dgrad_finish"):  # (1x1 / stride 1): dy + weights + dx read-modify-write + y + mask
        n, h, w, cin, cout = a[3:8]  # (the lattice form reads a quarter of dx as data; counted as the whole tensor)
        return 2.0 * (n * h * w * (3 * cin + cout) + cin * cout) + n * h * w * cin / 8.0
    n, h, w, cin, cout, k, s = a[3:10]
    big, small = n * h * w * cin, n * (h // s) * (w // s) * cout  # input-side / output-side activation elements
    # a strided 1x1 convolution only touches the sampled input pixels (fprop / wgrad read them; the dgrad still
    # writes the whole, mostly zero, gradient tensor)
    sampled = big // (s * s) if (k == 1 and s > 1 and ("dgrad" not in name or a[10] == 2)) else big
    total = sampled + small + cout * k * k * cin
    if name == "peclr_conv2d_dgrad" and a[10] == 1:
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


def dump_launch_table(prof, peak_tflops, peak_gbs, path):
    """Per conv shape (C-ABI call + geometry): launches, mean CUDA-event time, achieved TFLOP/s and algorithmic GB/s,
    and the fraction of that launch's own roofline max(FLOPs / tensor peak, bytes / HBM peak) -- the table the tuning
    decisions are read from (PECLR_BENCH_DUMP_LAUNCHES=<file>)."""
    import collections

    rows = collections.OrderedDict()
    for name, a, ms_k in prof:
        geo = tuple(a[3:6]) if name.startswith("peclr_stem") else tuple(a[3:8 if "finish" in name else 10])
        r = rows.setdefault((name.replace("peclr_", ""), geo), [0, 0.0, conv_flops(name, a), conv_bytes(name, a)])
        r[0] += 1
        r[1] += ms_k
    with open(path, "w") as f:
        f.write("%-22s %-34s %3s %9s %8s %8s %6s\n" % ("call", "N,H,W,Cin,Cout[,k,s]", "n", "us/launch", "TFLOP/s",
                                                       "GB/s", "frac"))
        for (name, geo), (n, ms_sum, fl, by) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            t = ms_sum / n * 1e-3
            bound = max(fl / (peak_tflops * 1e12), by / (peak_gbs * 1e9))
            f.write("%-22s %-34s %3d %9.1f %8.0f %8.0f %6.2f\n" % (name, ",".join(map(str, geo)), n, t * 1e6,
                                                                   fl / t / 1e12, by / t / 1e9, bound / t))


def conv_flops(name, a):
    """2*M*N*K of one conv launch from its C-ABI arguments."""
    if name.startswith("peclr_stem"):
        n, h, w = a[3], a[4], a[5]
        return 2.0 * n * (h // 2) * (w // 2) * 64 * 147
    if name.startswith("peclr_conv2d_dgrad_finish"):
        n, h, w, cin, cout = a[3:8]
        return 2.0 * n * h * w * cout * cin
    n, h, w, cin, cout, k, s = a[3:10]
    return 2.0 * n * (h // s) * (w // s) * cout * cin * k * k


# ---------------------------------------------------------------------------------------------- CPU arms
def _cpu_model(args, batch):
    """The reference's own Hybrid2Model (imported from /root/reference through oracle/ref_shims.py) where the reference
    tree exists -- this build container --, else its oracle port (the GPU box has no /root/reference)."""
    import torch

    from oracle import peclr_oracle as po
    from oracle import ref_shims

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = po.default_config(resnet_size=args.model, batch_size=batch, num_samples=batch * 1000,
                            num_of_mini_batch=args.accumulate)
    torch.manual_seed(5)
    kind = "port"
    model = None
    if ref_shims.reference_available() and not os.environ.get("PECLR_BENCH_FORCE_PORT"):
        try:
            ref = ref_shims.load_reference()
            model = ref.Hybrid2Model(ref.EasyDict(dict(cfg)))
            kind = "reference"
        except Exception:  # pragma: no cover  (fall back to the port rather than lose the bench line)
            model = None
    if model is None:
        model = po.OracleHybrid2Model(cfg)
    model.trainer = po._TrainerStub(world_size=1, max_epochs=100)
    model.setup("fit")
    (opt,), (sch,) = model.configure_optimizers()
    return po, model, opt, sch["scheduler"], kind


def _time_cpu_steps(args, batch, warm, steps):
    po, model, opt, sched, kind = _cpu_model(args, batch)
    data = po.synthetic_batch(batch, args.size, seed=5, structured=False)
    for _ in range(warm):
        po.oracle_step(model, data, opt, sched)
    t0 = time.perf_counter()
    for _ in range(steps):
        po.oracle_step(model, data, opt, sched)
    return (time.perf_counter() - t0) / steps, kind


def cpu_sample_pairs(args):
    """Both CPU legs use the SAME per-step sample: min(B, 32) pairs of the B-pair batch (a full ResNet-50 B = 128 step
    at 224^2 takes ~35 s on 8 cores: SURVEY 6), throughput scaled per image."""
    return min(args.batch, CPU_SAMPLE_PAIRS)


def _cpu_what(kind):
    return ("the reference's Hybrid2Model.training_step + backward + LARSWrapper(Adam).step imported from "
            "/root/reference" if kind == "reference" else
            "oracle port of the reference's Hybrid2Model.training_step + backward + LARSWrapper(Adam).step")


def cpu_baseline(args, budget_s=20.0):
    b = cpu_sample_pairs(args)
    dt, kind = _time_cpu_steps(args, b, 1, 2)
    return {"value": round(2 * b / dt, 2), "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
            "sample": "%s (fp32, torch CPU, %d threads): 1 warm-up + 2 timed steps at B=%d of %d pairs, %dx%d, "
                      "scaled per image" % (_cpu_what(kind), os.cpu_count(), b, args.batch, args.size, args.size)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = cpu_sample_pairs(args)
    po, model, opt, sched, kind = _cpu_model(args, b)
    data = po.synthetic_batch(b, args.size, seed=5, structured=False)
    # bounded run: at most ~3 minutes of CPU work whatever --steps / --warmup ask for (the count actually timed is
    # reported in "steps")
    t_probe = time.perf_counter()
    po.oracle_step(model, data, opt, sched)
    probe = time.perf_counter() - t_probe
    budget = max(1, int(150.0 / max(probe, 1e-3)))
    warm = min(args.warmup, max(0, budget // 4))
    steps = max(1, min(args.steps, budget - warm))
    for _ in range(warm):
        po.oracle_step(model, data, opt, sched)
    t0 = time.perf_counter()
    for _ in range(steps):
        po.oracle_step(model, data, opt, sched)
    dt = (time.perf_counter() - t0) / steps
    value = round(2 * b / dt, 2)
    sample = ("%s (fp32, torch CPU, %d threads); each step = B=%d pairs of the %d-pair batch, %dx%d, throughput per "
              "image; %d of the requested %d steps timed (bounded run)"
              % (_cpu_what(kind), os.cpu_count(), b, args.batch, args.size, args.size, steps, args.steps))
    print(json.dumps({
        "impl": "reference",
        "metric": "two-view images/sec, PeCLR pre-training step (fwd+bwd+LARS-Adam), ResNet-%s bs%d %dx%d"
                  % (args.model, args.batch, args.size, args.size),
        "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-%s PeCLR step on host cores, bounded sample B=%d pairs per step of the %d-pair "
                               "batch, %dx%d" % (args.model, b, args.batch, args.size, args.size),
                   "sample_pairs_per_step": b},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": kind, "sample": sample},
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
    ap.add_argument("--accumulate", type=int, default=1,
                    help="accumulate_grad_batches: micro-steps per optimiser step (BASELINE config 5: 16)")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--e2e-input", dest="e2e_input", default="fp32", choices=["u8", "fp32"],
                    help="host-side input of the end-to-end measurement: two ready-made normalised fp32 views per "
                         "sample (default: the reference's batch dict as its DataLoader hands it over), or raw uint8 "
                         "images + joints through the GPU augmentation (8x fewer bytes over PCIe, but the augmentation "
                         "kernel is extra work: measured 1.2 % slower end to end on one GPU)")
    ap.add_argument("--no-secondary", dest="no_secondary", action="store_true",
                    help="skip the ResNet-152 blocks (configs 4 / 5) the default workload also measures")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true",
                    help="skip the multi-GPU parity checks run before the timed region at world > 1")
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
