from peclr_b200.port_model import peclr_to_torchvision  # noqa: F401
