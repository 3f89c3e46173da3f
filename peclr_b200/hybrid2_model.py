"""PeCLR model (mirror of src/models/unsupervised/hybrid2_model.py:16-106): SimCLR whose projections are
moved back through the inverse crop-translation and inverse rotation before NT-Xent.  The whole chain after
the projection head -- statistics, both normalisations, translate, rotate, NT-Xent and its backward -- is one
CUDA launch (csrc/ntxent.cu)."""
from typing import Dict

from torch import Tensor

from .easydict import EasyDict as edict
from .ops import STAT_NAMES
from .simclr_model import SimCLR


class Hybrid2Model(SimCLR):
    uses_equivariance = True

    def __init__(self, config: edict):
        super().__init__(config)

    def contrastive_step(self, batch: Dict[str, Tensor]) -> Tensor:
        loss, stats = self._run_step(batch, want_stats=True)
        # the 16 detached projection statistics of get_projection_stats (hybrid2_model.py:92-106)
        self.train_metrics = {**self.train_metrics, **{name: stats[i] for i, name in enumerate(STAT_NAMES)}}
        return loss
