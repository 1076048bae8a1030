"""Negative (q-)expected improvement (reference: src/losses/negative_ei.py)."""
import numpy as np

from ...engine import acq_params
from .myopic_loss import myopic_loss


class negative_ei(myopic_loss):
    kind = "ei"

    def __init__(self, level_fn=None, **kwargs):
        super().__init__(**kwargs)
        self.level_fn = level_fn      # callable(outputs_old=...) -> level; default min(outputs_old) (:31-34)

    def acq_params(self, outputs_old=None, levels=None, level_fn=None, **kwargs):
        fn = self.level_fn if level_fn is None else level_fn
        if levels is None and fn is not None:
            levels = fn(outputs_old=outputs_old)
        return acq_params("ei", level=None if levels is None else float(np.squeeze(levels)))
