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
