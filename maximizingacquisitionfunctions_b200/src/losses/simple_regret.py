"""(q-)simple regret: E[min_i y_i]; q == 1 is the posterior mean (reference: src/losses/simple_regret.py)."""
from ...engine import acq_params
from .myopic_loss import myopic_loss


class simple_regret(myopic_loss):
    kind = "sr"

    def acq_params(self, **kwargs):
        return acq_params("sr")
