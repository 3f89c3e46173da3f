"""Host-side mirror of the reference interface, no GPU: module / state_dict layout, optimiser groups, schedule,
flag surface and config merging, synthetic batch schema, and that the step refuses to run without CUDA."""
import os

import numpy as np
import pytest
import torch

from oracle import peclr_oracle as po


@pytest.fixture(scope="module")
def model():
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    cfg = EasyDict(dict(po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64)))
    torch.manual_seed(0)
    return Hybrid2Model(cfg)


def test_state_dict_layout_matches_reference_golden(model, golden_dir):
    gold = np.load(os.path.join(golden_dir, "ckpt_layout.npz"))
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in gold["rn50_keys"]]
    assert ["x".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in gold["rn50_shapes"]]


def test_same_seed_gives_reference_initialisation(model):
    torch.manual_seed(0)
    oracle = po.OracleHybrid2Model(po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64))
    for (k1, v1), (k2, v2) in zip(oracle.state_dict().items(), model.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1


def test_parameters_are_views_of_one_flat_buffer(model):
    eng = model.engine
    assert eng.total == 24623680 and len(eng.segs) == 164  # SURVEY 8(a)-A11 (final_layer excluded: never trained)
    w = model.encoder.features[0].weight
    assert w.shape == (64, 3, 7, 7) and w.is_contiguous(memory_format=torch.channels_last)
    assert w.data_ptr() == eng.flat.data_ptr() and w.grad.data_ptr() == eng.grads.data_ptr()
    with torch.no_grad():
        w.view(-1)[0] if w.is_contiguous() else None
        model.projection_head[3].weight.fill_(0.25)
    s = eng._seg(model.projection_head[3])
    assert float(eng.flat[s.begin]) == 0.25


def test_optimizer_groups_and_schedule(model):
    class T:
        world_size, max_epochs = 1, 100

    model.trainer = T()
    model.setup("fit")
    assert model.train_iters_per_epoch == 64
    (opt,), (sch,) = model.configure_optimizers()
    # exclude_from_wt_decay quirk of the reference: only names containing "bias"/"bn" are excluded
    assert [len(g["params"]) for g in opt.param_groups] == [62, 104]
    assert opt.param_groups[0]["weight_decay"] == 1e-6 and opt.param_groups[1]["weight_decay"] == 0.0
    assert sch["interval"] == "step" and sch["frequency"] == 1
    base = 1e-4 * (8 * 1) ** 0.5
    s = sch["scheduler"]
    for step in range(700):
        want = po.warmup_cosine_lr(step, base, 10 * 64, 100 * 64)
        assert opt.param_groups[0]["lr"] == pytest.approx(want, rel=1e-9, abs=1e-15)
        s.last_epoch += 1  # advance without an optimizer.step() (no CUDA here)
        for g, lr in zip(opt.param_groups, s.get_lr()):
            g["lr"] = lr
    with pytest.raises(Exception):
        opt.step()  # the fused optimiser needs the CUDA library path


def test_step_refuses_to_run_on_cpu(model):
    from peclr_b200._lib import PeclrKernelError

    batch = po.synthetic_batch(2, 64, seed=1)
    with pytest.raises(PeclrKernelError):
        model.training_step(batch, 0)


def test_synthetic_batch_schema_equals_oracle_generator():
    from peclr_b200.synthetic import SyntheticTwoViewDataset, synthetic_batch

    a, b = synthetic_batch(4, 32, seed=9), po.synthetic_batch(4, 32, seed=9)
    assert set(a) == set(b) == {"transformed_image1", "transformed_image2", "angle_1", "angle_2", "jitter_x_1",
                                "jitter_x_2", "jitter_y_1", "jitter_y_2"}
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    assert a["angle_1"].dtype == torch.float64 and a["jitter_x_1"].dtype == torch.int64
    assert float(a["angle_1"].min()) >= -45 and float(a["jitter_x_2"].max()) <= 0 and float(a["jitter_y_1"].min()) >= -14
    item = SyntheticTwoViewDataset(10, 32, rotate=False)[3]
    assert "angle_1" not in item and item["transformed_image1"].shape == (3, 32, 32)


def test_flag_surface_and_config_merge():
    from peclr_b200.easydict import EasyDict
    from peclr_b200.experiments_utils import (get_general_args, get_model, update_model_params, update_train_params)
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.peclr_training import HYBRID2_CONFIG, TRAINING_CONFIG_PATH, read_json

    args = get_general_args("x", ["--rotate", "--crop", "--color_jitter", "-resnet_size", "152", "-epochs", "100",
                                  "-batch_size", "64", "-accumulate_grad_batches", "16", "-save_top_k", "1",
                                  "-save_period", "1", "-num_workers", "8", "-train_ratio", "0.9"])
    train = update_train_params(args, EasyDict(read_json(TRAINING_CONFIG_PATH)))
    assert train.batch_size == 64 and train.epochs == 100 and train.accumulate_grad_batches == 16
    assert train.augmentation_flags.rotate and train.augmentation_flags.crop and not train.augmentation_flags.resize
    assert train.train_ratio == pytest.approx(0.9) and train.seed == 5
    mp = update_model_params(EasyDict(read_json(HYBRID2_CONFIG)), args, 1234, train)
    assert mp.resnet_size == "152" and mp.num_samples == 1234 and mp.batch_size == 64 and mp.num_of_mini_batch == 16
    assert mp.optimizer == "LARS" and mp.lr == 1e-4  # not given on the command line -> JSON defaults stay
    assert get_model("hybrid2", False, False) is Hybrid2Model
    assert get_general_args("x", []).resnet_size == "18"  # reference default (experiments/utils.py:147-152)


def test_src_compat_imports():
    from src.experiments.utils import get_model
    from src.models.port_model import peclr_to_torchvision
    from src.models.unsupervised.hybrid2_model import Hybrid2Model

    assert get_model("hybrid2", False, False) is Hybrid2Model and callable(peclr_to_torchvision)


def test_experiment_naming_and_checkpoint_helpers(model, tmp_path, monkeypatch):
    """prepare_name / save_experiment_key / get_checkpoints / restore_model (src/experiments/utils.py:335-561)."""
    from peclr_b200.easydict import EasyDict
    from src.experiments.utils import get_checkpoints, prepare_name, restore_model, save_experiment_key

    flags = {"color_drop": False, "color_jitter": True, "crop": True, "cut_out": False, "gaussian_blur": False,
             "random_crop": False, "resize": True, "rotate": True, "gaussian_noise": False, "sobel_filter": False}
    train = EasyDict(batch_size=128, augmentation_flags=flags)
    assert prepare_name("hybrid2_", train, hybrid_naming=False) == "hybrid2_128C_CJ_Re_Ro"  # sorted codes
    both = EasyDict(batch_size=64, pairwise={"augmentation_flags": {"crop": True, "rotate": True}},
                    contrastive={"augmentation_flags": {"color_jitter": True}})
    assert prepare_name("hybrid1_", both, hybrid_naming=True) == "hybrid1_64_rel_C_Ro_con_CJ"

    monkeypatch.setenv("SAVED_META_INFO_PATH", str(tmp_path))
    monkeypatch.setenv("SAVED_MODELS_BASE_PATH", str(tmp_path))
    save_experiment_key("hybrid2_128C_Ro", "abc123", "meta.csv")
    save_experiment_key("hybrid2_128C_Ro", "def456", "meta.csv")
    assert (tmp_path / "meta.csv").read_text() == "hybrid2_128C_Ro,abc123\nhybrid2_128C_Ro,def456\n"

    ckpt_dir = tmp_path / "abc123" / "checkpoints"
    ckpt_dir.mkdir(parents=True)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["projection_head.3.weight"] = torch.full_like(sd["projection_head.3.weight"], 0.5)
    for epoch in (2, 10):
        torch.save({"state_dict": sd, "epoch": epoch}, ckpt_dir / f"epoch={epoch}.ckpt")
    assert get_checkpoints("abc123", number=1) == ["epoch=2.ckpt"]  # reverse lexicographic, as in the reference
    restored = restore_model(model, "abc123")  # newest by the integer in the name: epoch=10
    assert restored is model and float(model.projection_head[3].weight.flatten()[0]) == 0.5
    assert float(model.engine.flat[model.engine._seg(model.projection_head[3]).begin]) == 0.5  # flat buffer followed


def test_synthetic_train_val_split():
    from peclr_b200.synthetic import SyntheticTwoViewDataset, get_train_val_split

    data = SyntheticTwoViewDataset(100, 16, seed=5, train_ratio=0.9)
    train, val = get_train_val_split(data, batch_size=8, num_workers=0, drop_last=True)
    assert len(train.dataset) == 90 and len(val.dataset) == 10 and len(train) == 11 and len(val) == 1
    vb = next(iter(val))
    first_val = data.__class__(100, 16, seed=5, train_ratio=0.0)  # everything is validation: index 90 == val item 0
    first_val.is_training(False)
    assert torch.equal(vb["transformed_image1"][0], first_val[90]["transformed_image1"])
    assert vb["angle_1"].dtype == torch.float64 and vb["jitter_x_1"].dtype == torch.int64
    # the reference's default ratio (training_config.json: 0.9999999999) leaves no validation batch
    _, val = get_train_val_split(SyntheticTwoViewDataset(256, 16, train_ratio=0.9999999999), batch_size=8,
                                 num_workers=0, drop_last=True)
    assert len(val) == 0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = oracle port on the host cores): one JSON line with the
    contract's keys; under torchrun only rank 0 works and prints."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--batch", "4", "--size", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 2 and line["steps"] == 1 and line["warmup"] == 0 and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["dtype"] == "f32" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    # the reference itself where /root/reference exists (this container), its oracle port elsewhere (GPU box)
    from oracle import ref_shims

    assert cb["kind"] == ("reference" if ref_shims.reference_available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["config"]["sample_pairs_per_step"] == 4
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""  # other ranks exit without work


def test_bench_roofline_arithmetic():
    """bench.py's per-launch FLOP / byte accounting: 2MNK of a conv launch, every operand once."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fprop = ("peclr_conv2d_fprop", (0, 0, 0, 256, 56, 56, 64, 256, 1, 1, 0, 0, 0))
    assert bench.conv_flops(*fprop) == 2.0 * 256 * 56 * 56 * 256 * 64
    assert bench.conv_bytes(*fprop) == 2.0 * (256 * 56 * 56 * (64 + 256) + 256 * 64)
    acc = ("peclr_conv2d_dgrad", (0, 0, 0, 256, 56, 56, 256, 64, 1, 1, 1, 0))
    plain = ("peclr_conv2d_dgrad", (0, 0, 0, 256, 56, 56, 256, 64, 1, 1, 0, 0))
    assert bench.conv_bytes(*acc) - bench.conv_bytes(*plain) == 2.0 * 256 * 56 * 56 * 256  # read-modify-write of dx
    # the finishing dgrad (1x1): same FLOPs as the plain one; bytes = dy + weights + dx in and out + y + mask bits
    fin = ("peclr_conv2d_dgrad_finish", (0, 0, 0, 256, 56, 56, 256, 64, 0, 0, 0, 0))
    assert bench.conv_flops(*fin) == bench.conv_flops(*plain) and "peclr_conv2d_dgrad_finish" in bench.CONV_CALLS
    assert bench.conv_bytes(*fin) == 2.0 * (256 * 56 * 56 * (3 * 256 + 64) + 256 * 64) + 256 * 56 * 56 * 256 / 8
    stem = ("peclr_stem_fprop", (0, 0, 0, 256, 224, 224, 0, 0, 0))
    assert bench.conv_flops(*stem) == 2.0 * 256 * 112 * 112 * 64 * 147
    # the whole ResNet-50 @ 224 figure of SURVEY 8(d) is 24.287 GFLOP per image: one 3x3 256->256 @14x14 is 0.2312 of it
    c3 = ("peclr_conv2d_fprop", (0, 0, 0, 1, 14, 14, 256, 256, 3, 1, 0, 0, 0))
    assert abs(bench.conv_flops(*c3) / 1e6 - 2 * 115.6) < 0.1  # table A2: 115.6 MMAC per image
    ds = ("peclr_conv2d_fprop", (0, 0, 0, 256, 56, 56, 256, 512, 1, 2))  # strided 1x1: only the sampled pixels are read
    assert bench.conv_bytes(*ds) == 2.0 * (256 * 28 * 28 * (256 + 512) + 512 * 256)
    rec = [fprop + (0.10,), acc + (0.16,), ("peclr_conv2d_wgrad", (0,) * 11, 0.5)]
    out = bench.launch_bound_fraction(rec, 1362.3, 6540.8)
    assert out["launches"] == 2 and out["hbm_bound_launches"] == 2 and 0 < out["frac"] < 1
    assert bench.F_TRAIN[("50", 224)] == 24.287e9 and bench.F_TRAIN[("152", 224)] == 68.834e9


def test_checkpoint_save_and_trainer_restore(model, tmp_path):
    """ModelCheckpoint's file (Lightning layout) -> Trainer.restore: weights, Adam moments / step count, scheduler
    position and the epoch to continue with (reference: UpdatedModelCheckpoint + PL resume)."""
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model
    from peclr_b200.lightning import ModelCheckpoint, Trainer

    def prepared(m):
        tr = Trainer(max_epochs=100, default_root_dir=str(tmp_path))
        m.trainer = tr
        m.setup("fit")
        tr.optimizers, tr.lr_schedulers = m.configure_optimizers()
        return tr

    tr = prepared(model)
    opt, sch = tr.optimizers[0], tr.lr_schedulers[0]["scheduler"]
    eng = model.engine
    eng.exp_avg, eng.exp_avg_sq = torch.full_like(eng.flat, 0.125), torch.full_like(eng.flat, 0.5)
    opt.step_count = 7
    for _ in range(5):
        sch.step()
    tr.current_epoch, tr.global_step = 3, 7
    path = str(tmp_path / "checkpoints" / "epoch=3.ckpt")
    ModelCheckpoint()._save_model(path, tr, model)
    ckpt = torch.load(path, map_location="cpu")
    assert set(ckpt) >= {"epoch", "global_step", "state_dict", "optimizer_states", "lr_schedulers"}
    assert list(ckpt["state_dict"].keys()) == list(model.state_dict().keys())

    torch.manual_seed(1)
    other = Hybrid2Model(EasyDict(dict(po.default_config(resnet_size="50", batch_size=8, num_samples=8 * 64))))
    tr2 = prepared(other)
    assert tr2.restore(path, other) == 4 and tr2.global_step == 7
    for (k, a), b in zip(model.state_dict().items(), other.state_dict().values()):
        assert torch.equal(a, b), k
    assert tr2.optimizers[0].step_count == 7
    assert float(other.engine.exp_avg[0]) == 0.125 and float(other.engine.exp_avg_sq[-1]) == 0.5
    assert tr2.lr_schedulers[0]["scheduler"].last_epoch == sch.last_epoch == 5
    assert tr2.optimizers[0].param_groups[0]["lr"] == pytest.approx(opt.param_groups[0]["lr"])


def test_pretrained_without_weights_is_loud(monkeypatch, tmp_path, capsys):
    """The reference hard-codes pretrained=True (base_model.py:23): a missing ImageNet file must not pass silently."""
    from peclr_b200.easydict import EasyDict
    from peclr_b200.model_utils import get_wrapper_model

    cfg = EasyDict(dict(po.default_config(resnet_size="18", batch_size=2, num_samples=4)))
    monkeypatch.setenv("PECLR_PRETRAINED_DIR", str(tmp_path))
    monkeypatch.delenv("PECLR_ALLOW_RANDOM_INIT", raising=False)
    monkeypatch.delenv("PECLR_REQUIRE_PRETRAINED", raising=False)
    with pytest.warns(RuntimeWarning, match="RANDOM initialisation"):
        enc = get_wrapper_model(cfg, pretrained=True)
    assert "random" in enc.init_source and "RANDOM initialisation" in capsys.readouterr().err
    monkeypatch.setenv("PECLR_REQUIRE_PRETRAINED", "1")
    with pytest.raises(FileNotFoundError):
        get_wrapper_model(cfg, pretrained=True)
    # with the file in place the weights are loaded and the source recorded
    import torchvision

    monkeypatch.delenv("PECLR_REQUIRE_PRETRAINED")
    tv = torchvision.models.resnet18(weights=None)
    torch.save(tv.state_dict(), tmp_path / "resnet18.pth")
    enc = get_wrapper_model(cfg, pretrained=True)
    assert enc.init_source.startswith("imagenet:")
    assert torch.equal(enc.features[0].weight, tv.conv1.weight)


def test_trainer_steps_on_the_last_batch_of_a_partial_window():
    """Lightning 1.0.8 applies the accumulated gradient on the final batch of an epoch even if the accumulation
    window is not full: 5 batches at accumulate_grad_batches = 2 -> optimiser steps after batches 2, 4 and 5."""
    from peclr_b200.lightning import Trainer

    flags = [(i, last) for i, _, last in Trainer._with_last_flag(iter(range(5)), None)]
    assert flags == [(0, False), (1, False), (2, False), (3, False), (4, True)]
    acc = 2
    assert [i + 1 for i, last in flags if (i + 1) % acc == 0 or last] == [2, 4, 5]
    assert [(i, last) for i, _, last in Trainer._with_last_flag(iter(range(5)), 3)] == [(0, False), (1, False), (2, True)]
    assert list(Trainer._with_last_flag(iter(()), None)) == []


def test_model_checkpoint_restores_top_k_bookkeeping(tmp_path):
    from peclr_b200.lightning import ModelCheckpoint

    d = tmp_path / "checkpoints"
    d.mkdir()
    for e in (0, 3):
        (d / f"epoch={e}.ckpt").write_bytes(b"x")
    cb = ModelCheckpoint(save_top_k=2, dirpath=str(d))
    rec = {"callbacks": {"ModelCheckpoint": {"best": [(0.5, str(d / "epoch=3.ckpt")), (0.9, str(d / "epoch=0.ckpt")),
                                                      (0.7, str(d / "epoch=1.ckpt"))]}}}
    cb.restore_state(rec, str(d / "epoch=3.ckpt"))
    assert cb.best == [(0.5, str(d / "epoch=3.ckpt")), (0.9, str(d / "epoch=0.ckpt"))]  # the deleted file is dropped
    cb2 = ModelCheckpoint(save_top_k=2, dirpath=str(d))
    cb2.restore_state({}, str(d / "epoch=3.ckpt"))  # a checkpoint without the record (e.g. the reference's)
    assert [p for _, p in cb2.best] == [str(d / "epoch=0.ckpt"), str(d / "epoch=3.ckpt")]


def test_batched_augmentation_parameters_equal_the_scalar_specification():
    """draw_batch_params (vectorised, what GpuTwoViewAugmenter uses per step) == draw_view_params (the scalar
    restatement of SampleAugmenter's parameter drawing, pinned to the reference by the golden file): same random stream,
    same integers, bit-identical inverted warp matrices."""
    import random

    from peclr_b200.gpu_augment import GpuTwoViewAugmenter, draw_batch_params, draw_view_params, invert_affine

    rs = np.random.RandomState(0)
    b = 96
    joints = (rs.randn(b, 21, 3) * rs.uniform(5, 45, (b, 1, 1)) + rs.uniform(30, 200, (b, 1, 3))).astype(np.float32)
    for flags in (dict(rotate=True, crop=True, random_crop=True, resize=True, color_jitter=True),
                  dict(rotate=False, crop=True, random_crop=False, resize=True, color_jitter=False),
                  dict(rotate=True, crop=False, random_crop=True, resize=True, color_jitter=True)):
        random.seed(3)
        scalar = [(draw_view_params(x, (224, 240), flags), draw_view_params(x, (224, 240), flags)) for x in joints]
        random.seed(3)
        bp = draw_batch_params(joints, (224, 240), flags)
        for i in range(b):
            for v in (0, 1):
                d = scalar[i][v]
                for k in ("ox", "oy", "cw", "ch", "side", "jitter_x", "jitter_y", "crop_margin_scale"):
                    assert d[k] == bp[k][v, i], (k, i, v)
                if flags["rotate"]:
                    assert d["angle"] == bp["angle"][v, i]
                    assert np.array_equal(invert_affine(d["m_fwd"]), bp["m_inv"][v, i])
                if flags["color_jitter"]:
                    assert [d["h"], d["s"], d["a"], d["b"]] == list(bp["hsab"][v, i])
        aug = GpuTwoViewAugmenter(flags, device="cpu")
        assert np.array_equal(aug.table(scalar, b, 224, 240), aug.batch_table(bp, b, 224, 240))
    with pytest.raises(NotImplementedError):
        GpuTwoViewAugmenter(dict(resize=True, gaussian_blur=True))


def test_stage_ranges_partition_the_flat_gradient_buffer(model):
    """The per-stage all-reduce sends four contiguous ranges [stem + layer1 | layer2 | layer3 | layer4 + head] of the flat
    gradient buffer, in the order backward finishes them (layer4 first): together they cover every trained scalar
    exactly once and each boundary is a parameter boundary."""
    ranges = model._stage_ranges()
    assert set(ranges) == {3, 2, 1, -1}
    spans = sorted(ranges.values())
    assert spans[0][0] == 0 and spans[-1][1] == model.engine.total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    begins = {s.begin for s in model.engine.segs}
    assert all(lo in begins for lo, _ in spans)
    by_name = {s.name: s for s in model.engine.segs}
    lo4, hi4 = ranges[3]
    assert lo4 == by_name["encoder.features.7.0.conv1.weight"].begin and hi4 == model.engine.total
    head = by_name["projection_head.0.weight"]
    assert lo4 <= head.begin < hi4  # the head's gradients are final before the trunk's backward starts
    assert ranges[-1] == (0, by_name["encoder.features.5.0.conv1.weight"].begin)
    model.enable_overlapped_sync(True)  # one process: stays off
    assert model._after_stage_hook is None


def test_build_digest_covers_sources_only(tmp_path, monkeypatch):
    """The build stamp keys bench.py's `roofline.traffic` to the kernel sources: objects, the library and the stamp
    itself must not change it (it used to, so every rebuild invalidated the committed ncu capture)."""
    from peclr_b200 import build

    before = build._digest()
    junk = [os.path.join(build.CSRC, n) for n in ("zz_test_object.o", ".zz_test_stamp")]
    try:
        for j in junk:
            with open(j, "w") as f:
                f.write("not a source")
        assert build._digest() == before
    finally:
        for j in junk:
            if os.path.exists(j):
                os.remove(j)
    probe = os.path.join(build.CSRC, "zz_test_probe.cuh")
    try:
        with open(probe, "w") as f:
            f.write("// a source file does count\n")
        assert build._digest() != before
    finally:
        os.remove(probe)
