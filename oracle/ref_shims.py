"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (dahiyaaneesh/peclr).

Imports the reference's own hot-path modules from ``/root/reference`` inside this
build container so the oracle restatement (``oracle/peclr_oracle.py``) can be pinned
against the code it restates and golden vectors can be generated
(``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may import this file.

Eight third-party packages the reference imports are not installed (no network):
pytorch_lightning, pl_bolts, easydict, kornia, comet_ml, yacs, matplotlib, skimage.
They are replaced by inert stand-ins in ``sys.modules``; the two that carry
arithmetic (pl_bolts ``LARSWrapper`` / ``LinearWarmupCosineAnnealingLR``, pinned at
pytorch-lightning-bolts==0.2.2 in the reference's requirements.txt:105) are bound to
the restatements in ``oracle/peclr_oracle.py`` (parity for those two is therefore
*unpinned*: the library source is not available offline).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PECLR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


class _AttrDict(dict):
    """Recursive attribute dict standing in for easydict.EasyDict."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    import torch.nn as nn

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import peclr_oracle as po

    if "easydict" not in sys.modules:
        _module("easydict", EasyDict=_AttrDict)
    if "kornia" not in sys.modules:
        _module("kornia")
    if "comet_ml" not in sys.modules:
        _module("comet_ml", Experiment=type("Experiment", (), {}))
    if "matplotlib" not in sys.modules:
        plt = _module("matplotlib.pyplot", Axes=type("Axes", (), {}))
        axes = _module("matplotlib.axes", Axes=plt.Axes)
        _module("matplotlib", pyplot=plt, axes=axes, use=lambda *a, **k: None)
    if "yacs" not in sys.modules:
        cfg = _module("yacs.config", load_cfg=lambda f: None)
        _module("yacs", config=cfg)
    if "pytorch_lightning" not in sys.modules:

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

        class _Callback:
            def __init__(self, *a, **k):
                pass

        light = _module("pytorch_lightning.core.lightning", LightningModule=LightningModule)
        core = _module("pytorch_lightning.core", lightning=light, LightningModule=LightningModule)
        comet = _module("pytorch_lightning.loggers.comet", CometLogger=type("CometLogger", (), {}))
        loggers = _module("pytorch_lightning.loggers", comet=comet, CometLogger=comet.CometLogger)
        mc = _module("pytorch_lightning.callbacks.model_checkpoint", ModelCheckpoint=_Callback)
        cbs = _module(
            "pytorch_lightning.callbacks",
            model_checkpoint=mc,
            ModelCheckpoint=_Callback,
            LearningRateMonitor=_Callback,
            Callback=_Callback,
        )
        _module(
            "pytorch_lightning",
            core=core,
            loggers=loggers,
            callbacks=cbs,
            LightningModule=LightningModule,
            Trainer=type("Trainer", (), {}),
            seed_everything=lambda s: None,
        )
    if "pl_bolts" not in sys.modules:
        lars = _module("pl_bolts.optimizers.lars_scheduling", LARSWrapper=po.LARSWrapper)
        sched = _module(
            "pl_bolts.optimizers.lr_scheduler",
            LinearWarmupCosineAnnealingLR=po.LinearWarmupCosineAnnealingLR,
        )
        opt = _module("pl_bolts.optimizers", lars_scheduling=lars, lr_scheduler=sched)
        _module("pl_bolts", optimizers=opt)
    for k in ("BASE_PATH",):
        os.environ.setdefault(k, REFERENCE_ROOT)
    for k in ("DATA_PATH", "SAVED_MODELS_BASE_PATH", "SAVED_META_INFO_PATH"):
        os.environ.setdefault(k, "/tmp/peclr_ref_" + k.lower())


_REF = None


def load_reference():
    """Returns a namespace with the reference's hot-path symbols (imported, not copied)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_shims()
    # the reference lives in a top-level package called ``src``; this repo also ships a
    # ``src`` compatibility package, so import the reference's under a clean sys.path
    # head and detach it again afterwards.
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    # (the reference's ``src`` has no __init__.py -> it is a namespace package and would lose to this repo's
    # regular ``src`` package wherever it sits on sys.path, so the repo root is taken off the path meanwhile)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [p for p in saved_path if os.path.abspath(p or os.getcwd()) != repo_root]
    try:
        import importlib

        mutils = importlib.import_module("src.models.utils")
        base = importlib.import_module("src.models.base_model")
        simclr = importlib.import_module("src.models.unsupervised.simclr_model")
        hybrid2 = importlib.import_module("src.models.unsupervised.hybrid2_model")
        port = importlib.import_module("src.models.port_model")
        # flag surface / config merging / experiment naming (src/experiments/utils.py); its import chain reaches the
        # data loaders, which still use the torch 1.7 module path ``torch.tensor``
        if "torch.tensor" not in sys.modules:
            import torch

            sys.modules["torch.tensor"] = _module("torch.tensor", Tensor=torch.Tensor)
        try:
            exp_utils = importlib.import_module("src.experiments.utils")
            augmenter = importlib.import_module("src.data_loader.sample_augmenter")
        except Exception:  # pragma: no cover  (optional: only the CLI / batch-contract pinning tests need them)
            exp_utils = augmenter = None
        # base_model.py:23 hard-codes pretrained=True (needs network) -> force False.
        orig = mutils.get_wrapper_model
        base.get_wrapper_model = lambda config, pretrained, wrapper=False: orig(config, False, wrapper)
        ns = types.SimpleNamespace(
            utils=mutils,
            base_model=base,
            simclr_model=simclr,
            hybrid2_model=hybrid2,
            port_model=port,
            experiments_utils=exp_utils,
            sample_augmenter=augmenter,
            Hybrid2Model=hybrid2.Hybrid2Model,
            SimCLR=simclr.SimCLR,
            vanila_contrastive_loss=mutils.vanila_contrastive_loss,
            rotate_encoding=mutils.rotate_encoding,
            translate_encodings=mutils.translate_encodings,
            get_rotation_2D_matrix=mutils.get_rotation_2D_matrix,
            peclr_to_torchvision=port.peclr_to_torchvision,
            EasyDict=_AttrDict,
        )
    finally:
        sys.path[:] = saved_path
        ref_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    _REF = ns
    return ns
