from peclr_b200.resnet_model import ResNetModel  # noqa: F401
