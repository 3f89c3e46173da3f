from peclr_b200.peclr_training import read_json  # noqa: F401
