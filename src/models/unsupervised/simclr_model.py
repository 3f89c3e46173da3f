from peclr_b200.simclr_model import SimCLR  # noqa: F401
