from peclr_b200.hybrid2_model import Hybrid2Model  # noqa: F401
