"""(q-)lower / upper confidence bound via folded-normal residuals mu -/+ |sqrt(pi beta / 2) L z|
(reference: src/losses/confidence_bound.py:27-116)."""
from ...engine import acq_params
from .myopic_loss import myopic_loss


class confidence_bound(myopic_loss):
    kind = "cb"

    def __init__(self, beta=2.0, lower=True, transform_samples=True, **kwargs):
        super().__init__(**kwargs)
        self.beta = beta
        self.lower = lower
        self.transform_samples = transform_samples

    def acq_params(self, beta=None, lower=None, **kwargs):
        return acq_params("cb", beta=self.beta if beta is None else beta, lower=self.lower if lower is None else lower)
