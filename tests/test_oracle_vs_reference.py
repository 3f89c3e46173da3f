"""Pins the oracle restatement against the executed reference (build container only)."""
import numpy as np
import pytest
import torch

from oracle import peclr_oracle as po

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_shims import load_reference

    return load_reference()


def test_full_step_bit_identical_to_reference(ref):
    cfg = po.default_config(resnet_size="50", batch_size=4, num_samples=4 * 64)
    torch.manual_seed(0)
    theirs = ref.Hybrid2Model(ref.EasyDict(dict(cfg)))
    torch.manual_seed(0)
    ours = po.OracleHybrid2Model(cfg)
    for (k1, v1), (k2, v2) in zip(theirs.state_dict().items(), ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    batch = po.synthetic_batch(4, 64, seed=7)
    theirs.train(), ours.train()
    o1 = theirs.training_step({k: v.clone() for k, v in batch.items()}, 0)
    o2 = ours.training_step({k: v.clone() for k, v in batch.items()}, 0)
    assert set(o1) == set(o2) and len(o1) == 17
    for k in o1:
        assert torch.equal(o1[k], o2[k]), k
    o1["loss"].backward(), o2["loss"].backward()
    g1, g2 = po.named_grads(theirs), po.named_grads(ours)
    assert list(g1) == list(g2) and len(g1) == 164
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k
    # optimiser + schedule through the reference's own configure_optimizers
    theirs.trainer = po._TrainerStub()
    theirs.setup("fit"), ours.setup("fit")
    (op1,), (sc1,) = theirs.configure_optimizers()
    (op2,), (sc2,) = ours.configure_optimizers()
    assert [len(g["params"]) for g in op1.param_groups] == [len(g["params"]) for g in op2.param_groups] == [62, 104]
    for _ in range(2):
        sc1["scheduler"].step(), sc2["scheduler"].step()
    op1.step(), op2.step()
    for (k, a), (_, b) in zip(theirs.named_parameters(), ours.named_parameters()):
        assert torch.equal(a, b), k


def test_ops_identical(ref):
    g = torch.Generator().manual_seed(11)
    enc = torch.randn(6, 64, 2, generator=g)
    ang = torch.floor(torch.rand(6, generator=g, dtype=torch.float64) * 90 - 45)
    assert torch.equal(ref.rotate_encoding(enc.clone(), ang), po.rotate_encoding(enc.clone(), ang))
    tx, ty = torch.rand(6, generator=g), torch.rand(6, generator=g)
    assert torch.equal(ref.translate_encodings(enc.clone(), tx, ty), po.translate_encodings(enc.clone(), tx, ty))
    z = torch.nn.functional.normalize(torch.randn(16, 128, generator=g))
    assert torch.equal(ref.vanila_contrastive_loss(z[:8], z[8:]), po.vanila_contrastive_loss(z[:8], z[8:]))


def test_cli_config_and_naming_identical_to_reference(ref, monkeypatch):
    """The reference's own get_general_args / update_train_params / update_model_params / prepare_name
    (src/experiments/utils.py:29-163,276-393,608-615), executed, against this build's mirror of them."""
    import json
    import os
    import sys

    from peclr_b200 import experiments_utils as ours
    from peclr_b200.easydict import EasyDict
    from peclr_b200.peclr_training import HYBRID2_CONFIG, TRAINING_CONFIG_PATH

    theirs = ref.experiments_utils
    assert theirs is not None, "src.experiments.utils of the reference did not import"
    ref_cfg = "/root/reference/src/experiments/config"
    def no_aug(d):  # (this build ships only the augmentation_params the synthetic generator uses: a subset)
        return {k: v for k, v in json.loads(json.dumps(d)).items() if k != "augmentation_params"}

    for mine, name in ((TRAINING_CONFIG_PATH, "training_config.json"), (HYBRID2_CONFIG, "hybrid2_config.json")):
        a, b = json.load(open(mine)), json.load(open(os.path.join(ref_cfg, name)))
        assert no_aug(a) == no_aug(b), name  # same shipped defaults
        for k, v in a.get("augmentation_params", {}).items():
            assert b["augmentation_params"][k] == v, k
    argvs = [
        [],
        ["--rotate", "--crop", "-resnet_size", "50", "-epochs", "100", "-batch_size", "128", "-accumulate_grad_batches",
         "16", "-save_top_k", "1", "-save_period", "1", "-num_workers", "8"],  # README.md:51 of the reference
        ["--color_jitter", "--random_crop", "--rotate", "--crop", "--resize", "-resnet_size", "152", "-sources",
         "freihand", "-sources", "youtube", "-lr", "0.001", "-optimizer", "adam", "-seed", "7", "-lr_max_epochs", "50",
         "-tag", "a", "-tag", "b", "-train_ratio", "0.9", "-log_interval", "step", "-meta_file", "m.csv"],
    ]
    for argv in argvs:
        monkeypatch.setattr(sys, "argv", ["peclr_training.py"] + argv)
        a_ref = vars(theirs.get_general_args("Hybrid model 2 training script."))
        a_our = vars(ours.get_general_args("Hybrid model 2 training script.", argv))
        for k, v in a_ref.items():  # every reference flag exists here with the same parsed value / default
            assert k in a_our and a_our[k] == v, (argv, k, v, a_our.get(k))
        t_ref = theirs.update_train_params(theirs.get_general_args("x"),
                                           ref.EasyDict(json.load(open(os.path.join(ref_cfg, "training_config.json")))))
        t_our = ours.update_train_params(ours.get_general_args("x", argv), EasyDict(json.load(open(TRAINING_CONFIG_PATH))))
        assert no_aug(t_ref) == no_aug(t_our), argv
        m_ref = theirs.update_model_params(ref.EasyDict(json.load(open(HYBRID2_CONFIG))), theirs.get_general_args("x"),
                                           4321, t_ref)
        m_our = ours.update_model_params(EasyDict(json.load(open(HYBRID2_CONFIG))), ours.get_general_args("x", argv),
                                         4321, t_our)
        assert json.loads(json.dumps(m_ref)) == json.loads(json.dumps(m_our)), argv
        assert theirs.prepare_name("hybrid2_", t_ref) == ours.prepare_name("hybrid2_", t_our)
    both = dict(batch_size=64, pairwise={"augmentation_flags": {"crop": True, "rotate": True, "flip": False}},
                contrastive={"augmentation_flags": {"color_jitter": True, "gaussian_blur": True}})
    assert theirs.prepare_name("hybrid1_", ref.EasyDict(both), True) == ours.prepare_name("hybrid1_", EasyDict(both), True)


def test_translate2_and_rotation_matrix_identical(ref):
    g = torch.Generator().manual_seed(12)
    enc = torch.randn(5, 64, 2, generator=g)
    tx, ty = torch.rand(5, generator=g), torch.rand(5, generator=g)
    want = ref.utils.translate_encodings2(enc.clone(), tx, ty)
    assert torch.equal(want, enc + torch.stack([tx, ty], dim=1)[:, None, :])  # what the CUDA op is tested against
    ang = torch.floor(torch.rand(5, generator=g, dtype=torch.float64) * 90 - 45)
    cx, cy = torch.randn(5, generator=g), torch.randn(5, generator=g)
    assert torch.equal(ref.get_rotation_2D_matrix(ang, cx, cy, 1.0), po.get_rotation_2D_matrix(ang, cx, cy, 1.0))


def test_batch_contract_matches_executed_augmenter(ref):
    """what the reference's SampleAugmenter really emits for the per-sample correction parameters
    (executed here on a synthetic image + joints with the reference's own training_config.json), against the
    synthetic two-view generator that replaces src/data_loader: angle = integer-valued python float in [-45, 45)
    -> float64 after collation, jitter_x / jitter_y = python int in {-14..0} -> int64."""
    import json
    import random

    from peclr_b200.synthetic import SyntheticTwoViewDataset, synthetic_batch

    assert ref.sample_augmenter is not None
    cfg = json.load(open("/root/reference/src/experiments/config/training_config.json"))
    flags = ref.EasyDict(cfg["augmentation_flags"])
    flags.crop = flags.rotate = flags.resize = True
    aug = ref.sample_augmenter.SampleAugmenter(flags, ref.EasyDict(cfg["augmentation_params"]))
    random.seed(0)
    g = torch.Generator().manual_seed(0)
    img = (np.random.RandomState(0).rand(224, 224, 3) * 255).astype(np.uint8)
    angles, jit = [], []
    for _ in range(300):
        joints = torch.rand(21, 3, generator=g) * 60 + 80  # a hand well inside the image: no clipping at the edge
        out_img, _, _ = aug.transform_sample(img, joints)
        assert out_img.shape == (128, 128, 3)  # resize_shape of the reference config
        angles.append(aug.angle)
        jit += [aug.jitter_x, aug.jitter_y]
    assert all(isinstance(a, float) and float(a).is_integer() for a in angles) and all(isinstance(j, int) for j in jit)
    assert -45 <= min(angles) and max(angles) <= 44 and min(jit) >= -14 and max(jit) <= 0
    collated = torch.utils.data.default_collate([{"angle_1": angles[0], "jitter_x_1": jit[0]},
                                                 {"angle_1": angles[1], "jitter_x_1": jit[2]}])
    assert collated["angle_1"].dtype == torch.float64 and collated["jitter_x_1"].dtype == torch.int64
    # the synthetic generator covers the same value sets with the same dtypes, batched and per sample
    b = synthetic_batch(512, 8, seed=1)
    for k in (1, 2):
        a = b[f"angle_{k}"]
        assert a.dtype == torch.float64 and torch.equal(a, a.floor()) and -45 <= a.min() and a.max() <= 44
        for ax in "xy":
            j = b[f"jitter_{ax}_{k}"]
            assert j.dtype == torch.int64 and -14 <= j.min() and j.max() <= 0
    assert set(angles) <= set(b["angle_1"].tolist()) | set(b["angle_2"].tolist()) | set(float(v) for v in range(-45, 45))
    item = torch.utils.data.default_collate([SyntheticTwoViewDataset(4, 8)[i] for i in range(2)])
    assert item["angle_1"].dtype == torch.float64 and item["jitter_y_2"].dtype == torch.int64
    assert item["transformed_image1"].shape == (2, 3, 8, 8) and item["transformed_image1"].dtype == torch.float32


def test_epoch_hooks_and_schedule_setup_identical(ref):
    """BaseModel.setup / training_epoch_end / validation_epoch_end / exclude_from_wt_decay (base_model.py:30-55,106-127)
    executed on the reference class against this build's BaseModel (host logic, CPU)."""
    from peclr_b200.easydict import EasyDict
    from peclr_b200.hybrid2_model import Hybrid2Model

    cfg = po.default_config(resnet_size="18", batch_size=16, num_samples=16 * 37 + 5, projection_head_input_dim=512)
    torch.manual_seed(0)
    theirs = ref.Hybrid2Model(ref.EasyDict(dict(cfg)))
    torch.manual_seed(0)
    ours = Hybrid2Model(EasyDict(dict(cfg)))
    logged = {}
    theirs.log = lambda name, value, **kw: logged.__setitem__(name, value)
    for world in (1, 2, 8):
        theirs.trainer = po._TrainerStub(world_size=world, max_epochs=50)
        ours.trainer = po._TrainerStub(world_size=world, max_epochs=50)
        theirs.setup("fit"), ours.setup("fit")
        assert theirs.train_iters_per_epoch == ours.train_iters_per_epoch == (16 * 37 + 5) // (16 * world)
    g = torch.Generator().manual_seed(3)
    outputs = [{"loss": torch.rand((), generator=g), "proj1x_mean": torch.randn((), generator=g)} for _ in range(7)]
    theirs.training_epoch_end(outputs), ours.training_epoch_end(outputs)
    assert set(theirs.train_metrics_epoch) == set(ours.train_metrics_epoch)
    for k in theirs.train_metrics_epoch:
        assert torch.equal(theirs.train_metrics_epoch[k], ours.train_metrics_epoch[k]), k
    assert torch.equal(logged["checkpoint_saving_loss"], ours._logged["checkpoint_saving_loss"])
    theirs.validation_epoch_end(outputs), ours.validation_epoch_end(outputs)
    assert torch.equal(theirs.validation_metrics_epoch["loss"], ours.validation_metrics_epoch["loss"])
    # weight-decay groups: same parameter NAMES in the decayed / excluded groups
    def names(model, groups):
        by_id = {id(p): n for n, p in model.named_parameters()}
        return [[by_id[id(p)] for p in grp["params"]] for grp in groups]

    g1 = theirs.exclude_from_wt_decay(theirs.named_parameters(), weight_decay=1e-6)
    g2 = ours.exclude_from_wt_decay(ours.named_parameters(), weight_decay=1e-6)
    assert names(theirs, g1) == names(ours, g2)
    assert [grp["weight_decay"] for grp in g1] == [grp["weight_decay"] for grp in g2] == [1e-6, 0.0]


def test_checkpoint_helpers_identical(ref, monkeypatch, tmp_path):
    """get_latest_checkpoint / get_encoder_state_dict (src/models/utils.py:189-225), get_checkpoints / restore_model
    (src/experiments/utils.py:535-561) executed on the same checkpoint directory as this build's versions."""
    import os

    from peclr_b200 import experiments_utils as eu
    from peclr_b200 import model_utils as mu

    base = ref.utils.SAVED_MODELS_BASE_PATH  # captured by the reference at import time
    monkeypatch.setenv("SAVED_MODELS_BASE_PATH", base)
    key = "pin_%d" % os.getpid()
    ckpt_dir = os.path.join(base, key, "checkpoints")
    os.makedirs(ckpt_dir, exist_ok=True)
    sd = {"encoder.features.0.weight": torch.randn(2, 3), "projection_head.0.bias": torch.randn(4),
          "encoder.final_layer.0.bias": torch.randn(3)}
    try:
        for epoch in (2, 10, 9):
            torch.save({"state_dict": {k: v + epoch for k, v in sd.items()}, "epoch": epoch},
                       os.path.join(ckpt_dir, f"epoch={epoch}.ckpt"))
        assert ref.utils.get_latest_checkpoint(key) == mu.get_latest_checkpoint(key)
        assert mu.get_latest_checkpoint(key).endswith("epoch=10.ckpt")  # by the integer, not lexicographically
        assert ref.utils.get_latest_checkpoint(key, "epoch=2.ckpt") == mu.get_latest_checkpoint(key, "epoch=2.ckpt")
        e1, e2 = ref.utils.get_encoder_state_dict(key, ""), mu.get_encoder_state_dict(key, "")
        assert list(e1) == list(e2) == ["features.0.weight", "final_layer.0.bias"]
        assert all(torch.equal(e1[k], e2[k]) for k in e1)
        assert ref.experiments_utils.get_checkpoints(key, 2) == eu.get_checkpoints(key, 2)
        m1, m2 = torch.nn.Linear(3, 2, bias=False), torch.nn.Linear(3, 2, bias=False)
        full = {"weight": torch.randn(2, 3)}
        torch.save({"state_dict": full}, os.path.join(ckpt_dir, "epoch=11.ckpt"))
        ref.experiments_utils.restore_model(m1, key), eu.restore_model(m2, key)
        assert torch.equal(m1.weight, m2.weight) and torch.equal(m2.weight.detach(), full["weight"])
    finally:
        for f in os.listdir(ckpt_dir):
            os.remove(os.path.join(ckpt_dir, f))
        os.removedirs(ckpt_dir)


@pytest.mark.reference
def test_rn25d_oracle_equals_reference():
    """The downstream network's oracle restatement (oracle/rn25d_oracle.py) against the reference module itself
    (src/models/rn_25D_wMLPref.py, imported from /root/reference: it needs only torch + torchvision): same state_dict
    for the same seed, bit-identical outputs in eval mode, with and without a per-sample camera matrix; and the
    product's holder module has the same keys and initial values."""
    import importlib.util

    from oracle.rn25d_oracle import OracleRN25D
    from peclr_b200.rn_25D_wMLPref import RN_25D_wMLPref

    spec = importlib.util.spec_from_file_location("ref_rn25d", "/root/reference/src/models/rn_25D_wMLPref.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    theirs = ref.RN_25D_wMLPref("rn50")
    torch.manual_seed(0)
    ours = OracleRN25D("rn50")
    torch.manual_seed(0)
    product = RN_25D_wMLPref("rn50")
    a, b, c = theirs.state_dict(), ours.state_dict(), product.state_dict()
    assert list(a) == list(b) == list(c) and len(a) == 336
    assert all(torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]) for k in a)
    product.load_state_dict(a)  # the released checkpoints' layout loads
    theirs.eval(), ours.eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 224, 224, generator=g)
    K = torch.tensor([[[400.0, 0.0, 110.0], [0.0, 395.0, 115.0], [0.0, 0.0, 1.0]]]).repeat(2, 1, 1)
    K[1, 0, 0] = 380.0
    with torch.no_grad():
        for kk in (None, K):
            ra, oa = theirs(x, kk), ours(x, kk)
            assert set(ra) == set(oa) == {"kp3d", "zrel", "kp2d", "kp25d"}
            for k in ra:
                assert torch.equal(ra[k], oa[k]), k
    with pytest.raises(Exception):
        RN_25D_wMLPref("rn18")
