from peclr_b200.base_model import BaseModel  # noqa: F401
