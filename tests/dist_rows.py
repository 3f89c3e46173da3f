"""Test helper: row bookkeeping of the data-parallel step: rank r holds B pairs; in the global 2*B*R batch NT-Xent sees,
view-1 rows of all ranks come first, then view-2 rows, so positives stay B*R apart.
The CUDA kernel uses the same mapping (global_row in csrc/ntxent.cu)."""
import numpy as np


def global_rows(rank: int, world: int, b: int) -> np.ndarray:
    """Global row indices of rank `rank`'s 2B local rows (view 1 then view 2)."""
    return np.concatenate([rank * b + np.arange(b), world * b + rank * b + np.arange(b)])


local_rows = global_rows
