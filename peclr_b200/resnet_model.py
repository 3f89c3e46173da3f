"""ResNet encoder with the reference's module / state_dict layout, executed by the sm_100a kernels.

Mirror of the reference's ``ResNetModel`` (src/models/resnet_model.py:6-56): ``features`` =
Sequential(conv1, bn1, relu, maxpool, layer1..4, AdaptiveAvgPool2d(1)) and an (unused in pre-training)
``final_layer`` = Sequential(Linear(in_features, 64)); ``forward`` returns the flattened pooled features in
``mode == "pretraining"``.  The modules below only *hold* parameters/buffers under the reference's names
(so ``state_dict()`` round-trips through ``peclr_to_torchvision``, port_model.py:7-48); the arithmetic is
done by ``peclr_b200.engine.TrunkEngine`` on NHWC bf16 activations.
"""
import torch
import torch.nn as nn

# torchvision.models.resnet{18..152}: block kind, blocks per layer
ARCH = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
    "resnet152": ("bottleneck", [3, 8, 36, 3]),
}


class ConvHolder(nn.Module):
    """Holds a (Cout, Cin, k, k) weight stored channels_last (= the kernels' [Cout][k*k][Cin] layout)."""

    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.cin, self.cout, self.k, self.stride = cin, cout, k, stride
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).contiguous(memory_format=torch.channels_last))

    def extra_repr(self):
        return f"{self.cin}, {self.cout}, kernel_size={self.k}, stride={self.stride}"


class BNHolder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.num_features = c
        self.eps, self.momentum = 1e-5, 0.1
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Placeholder(nn.Module):
    """Parameter-free stage (ReLU / MaxPool / AvgPool) kept so that Sequential indices match torchvision's."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class Block(nn.Module):
    """BasicBlock / Bottleneck parameter holder (registration order = torchvision's state_dict order)."""

    def __init__(self, kind, inplanes, planes, stride):
        super().__init__()
        self.kind = kind
        if kind == "basic":
            out = planes
            self.conv1, self.bn1 = ConvHolder(inplanes, planes, 3, stride), BNHolder(planes)
            self.conv2, self.bn2 = ConvHolder(planes, planes, 3, 1), BNHolder(planes)
        else:
            out = planes * 4
            self.conv1, self.bn1 = ConvHolder(inplanes, planes, 1, 1), BNHolder(planes)
            self.conv2, self.bn2 = ConvHolder(planes, planes, 3, stride), BNHolder(planes)  # v1.5: stride on 3x3
            self.conv3, self.bn3 = ConvHolder(planes, out, 1, 1), BNHolder(out)
        self.downsample = None
        if stride != 1 or inplanes != out:
            self.downsample = nn.Sequential(ConvHolder(inplanes, out, 1, stride), BNHolder(out))
        self.out_channels = out

    def convs(self):
        if self.kind == "basic":
            return [(self.conv1, self.bn1), (self.conv2, self.bn2)]
        return [(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)]


class ResNetModel(nn.Module):
    def __init__(self, config, mode=""):
        super().__init__()
        self.mode = mode
        name = config.model.backend_model.lower()
        if name not in ARCH:
            raise NotImplementedError(name)
        self.arch = name
        kind, counts = ARCH[name]
        layers, inplanes = [], 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), counts)):
            blocks = []
            for j in range(n):
                blk = Block(kind, inplanes, planes, (1 if i == 0 else 2) if j == 0 else 1)
                inplanes = blk.out_channels
                blocks.append(blk)
            layers.append(nn.Sequential(*blocks))
        self.in_features = inplanes
        self.features = nn.Sequential(
            ConvHolder(3, 64, 7, 2), BNHolder(64), Placeholder("ReLU"), Placeholder("MaxPool2d(3, 2, 1)"),
            *layers, Placeholder("AdaptiveAvgPool2d(1)"),
        )
        self._init_like_reference(bool(config.model.pretrained))  # also creates final_layer (RNG order)
        self.engine = None  # bound by the owning model (peclr_b200.base_model.BaseModel)

    def _init_like_reference(self, pretrained):
        """Same initial values (and RNG consumption) as the reference: a torchvision ResNet built on the CPU,
        then final_layer's default nn.Linear init (resnet_model.py:13-29).  ImageNet weights cannot be
        downloaded here (no network): `pretrained` loads them from $PECLR_PRETRAINED_DIR/<arch>.pth.  The reference's
        BaseModel hard-codes pretrained=True (base_model.py:23), i.e. it ALWAYS starts from ImageNet weights; when
        the file is missing this build says so loudly (a warning on stderr and through `warnings`) instead of
        silently training from a different starting point, or raises if PECLR_REQUIRE_PRETRAINED=1.  The source of
        the initial weights is kept in `self.init_source` (the trainer logs it and stores it in checkpoints)."""
        import os
        import sys
        import warnings

        import torchvision

        tv = getattr(torchvision.models, self.arch)(weights=None, norm_layer=nn.BatchNorm2d)
        self.init_source = "random (torchvision default init)"
        if pretrained:
            root = os.environ.get("PECLR_PRETRAINED_DIR", "")
            path = os.path.join(root, self.arch + ".pth")
            if os.path.isfile(path):
                tv.load_state_dict(torch.load(path, map_location="cpu"))
                self.init_source = "imagenet:" + path
            elif os.environ.get("PECLR_REQUIRE_PRETRAINED", "0") == "1":
                raise FileNotFoundError(
                    f"pretrained=True (as the reference hard-codes) but {path!r} does not exist; put the torchvision "
                    f"{self.arch} ImageNet state_dict there (set PECLR_PRETRAINED_DIR), or unset "
                    "PECLR_REQUIRE_PRETRAINED to accept random initialisation")
            elif os.environ.get("PECLR_ALLOW_RANDOM_INIT", "0") != "1":
                msg = (f"peclr_b200: pretrained=True but no ImageNet weights at {path!r} (PECLR_PRETRAINED_DIR="
                       f"{root!r}): the encoder starts from RANDOM initialisation, unlike the reference, which "
                       "downloads torchvision's ImageNet weights.  Set PECLR_ALLOW_RANDOM_INIT=1 to silence this, "
                       "PECLR_REQUIRE_PRETRAINED=1 to make it an error.")
                warnings.warn(msg, RuntimeWarning, stacklevel=3)
                print("WARNING: " + msg, file=sys.stderr)
        self.final_layer = nn.Sequential(nn.Linear(tv.fc.in_features, 21 * 3 + 1))
        tv_sd = [(k, v) for k, v in tv.state_dict().items() if not k.startswith("fc.")]
        own = self.features.state_dict()
        assert len(tv_sd) == len(own)
        with torch.no_grad():
            for (k_tv, v_tv), (k_own, v_own) in zip(tv_sd, own.items()):
                assert k_tv.split(".")[-1] == k_own.split(".")[-1] and v_tv.shape == v_own.shape, (k_tv, k_own)
                v_own.copy_(v_tv)

    def forward(self, x):
        """x: fp32 NCHW images (both views concatenated).  Runs the CUDA trunk; returns [n, in_features] fp32.
        Standalone use is inference-style (no autograd through the trunk): training goes through the
        owning model's step, which drives the same engine with its backward."""
        if self.engine is None:
            raise RuntimeError("ResNetModel is not bound to an engine; construct it through Hybrid2Model/SimCLR")
        z = self.engine.encode(x)
        if self.mode == "pretraining":
            return z
        z = self.final_layer(z)
        return z[:, : 21 * 3], None, z[:, -1]
