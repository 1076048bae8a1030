"""Negative (q-)probability of improvement; hard indicator or sigmoid relaxation
(reference: src/losses/negative_pi.py, utils.soft_less src/utils/math_ops.py:56-69)."""
import numpy as np

from ...engine import acq_params
from .myopic_loss import myopic_loss


class negative_pi(myopic_loss):
    kind = "pi"

    def __init__(self, level_fn=None, temperature=None, **kwargs):
        super().__init__(**kwargs)
        self.level_fn = level_fn
        self.temperature = temperature

    def acq_params(self, outputs_old=None, levels=None, level_fn=None, temperature=None, **kwargs):
        fn = self.level_fn if level_fn is None else level_fn
        if levels is None and fn is not None:
            levels = fn(outputs_old=outputs_old)
        if temperature is None:
            temperature = self.temperature
        return acq_params("pi", level=None if levels is None else float(np.squeeze(levels)), temperature=temperature)
