"""Base class of the four myopic acquisition losses (reference: src/losses/myopic_loss.py).

Everything is a loss to minimise.  ``evaluate`` dispatches like the reference (:84-90): a pool of
size one takes the closed form (with or without base samples), anything else the reparameterised
Monte-Carlo estimator with base samples ``fantasy_rvs[M, q]``.  Evaluation is eager: NumPy in,
NumPy out, computed by the CUDA library; ``value_and_grad`` is the counterpart of the
``tf.gradients`` pass the reference's optimisers add.
"""
import logging

import numpy as np

from ...engine import acq_params
from ..utils import atleast_pools
from .abstract_loss import abstract_loss

logger = logging.getLogger(__name__)


class myopic_loss(abstract_loss):
    kind = None          # 'ei' | 'pi' | 'sr' | 'cb'

    def __init__(self, model=None, num_fantasies=2 ** 10, num_rand_features=2 ** 10, sharding_config=None,
                 use_rand_features=False, cache_rvs=False, **kwargs):
        super().__init__(**kwargs)
        if use_rand_features:
            raise NotImplementedError("random-feature sample paths are outside the accelerated path (SURVEY 2, row 3)")
        self.model = model
        self.num_fantasies = num_fantasies
        self.num_rand_features = num_rand_features
        self.use_rand_features = use_rand_features
        self.cache_rvs = cache_rvs
        # kept for interface parity; restarts are independent so device-side chunking replaces sharded_fn
        self.sharding_config = self.update_dict({
            "closed_form": {"shard_limit": 2 ** 16, "exact_sharding": True},
            "monte_carlo": {"shard_limit": 2 ** 12, "exact_sharding": True},
        }, sharding_config)

    # ---- per-class acquisition parameters -----------------------------------------------------
    def acq_params(self, outputs_old=None, levels=None, **kwargs):
        raise NotImplementedError

    # ---- evaluation -----------------------------------------------------------------------------
    def evaluate(self, pools, inputs_old, outputs_old, parallelism=None, fantasy_rvs=None, subroutine=None,
                 model=None, num_pending=0, weights=None, incumbents=None, return_grad=False, **kwargs):
        """losses (+ extras) for pools [S, q, d].

        Returns ``(losses, (means, cov_or_var, chol))`` like myopic_loss.sharded_subroutine (:179):
        MC: losses [S], means [S,q,1], cov [S,q,q], chol [S,q,q];
        closed form: losses [S,1,1], means [S,1,1], var [S,1,1], chol None.
        With ``return_grad`` the gradient d(sum losses)/d(pools[:, num_pending:]) is appended.
        """
        if model is None:
            model = self.model
        pools = np.ascontiguousarray(atleast_pools(pools))
        S, q, d = pools.shape
        if parallelism is None:
            parallelism = q
        assert parallelism == q, "parallelism must equal the pool size"
        eng = model.bind(inputs_old, outputs_old)
        if np.ndim(inputs_old) == 4:
            # fantasy-padded observations: closed form averaged over the fantasies on the device
            # (the reference returns [F,S,1,1] and appraise_inputs takes the mean over axis 0, :227-229)
            assert q == 1, "fantasy-conditioned losses are closed-form (q == 1)"
            prm = self.acq_params(outputs_old=None, **kwargs)
            losses, grad = eng.fantasies_value_and_grad(prm, pools[:, 0, :], need_grad=return_grad)
            out = (losses.reshape(S, 1, 1), None)
            return out + (None if grad is None else grad.reshape(S, 1, d),) if return_grad else out
        prm = self.acq_params(outputs_old=outputs_old, **kwargs)
        # reference dispatch (:84-90): a pool of one takes the closed form whether or not base samples were
        # passed; the Monte-Carlo estimator at q == 1 has to be asked for (subroutine=self.monte_carlo)
        closed = (q == 1) if subroutine is None else (subroutine == self.closed_form)
        if closed:
            assert q == 1, "closed forms exist for q == 1 only"
            losses, grad = eng.value_and_grad(prm, pools, None, need_grad=return_grad)
            mu, var, _ = eng.last_extras(S, 1, want_chol=False)
            out = (losses.reshape(S, 1, 1), (mu, var.reshape(S, 1, 1), None))
        else:
            if fantasy_rvs is None:
                fantasy_rvs = self.rng.randn(self.num_fantasies, q)
            losses, grad = eng.value_and_grad(prm, pools, np.asarray(fantasy_rvs, dtype=np.float64),
                                              p_pending=num_pending, weights=weights, incumbents=incumbents,
                                              need_grad=return_grad)
            out = (losses, eng.last_extras(S, q))
        return out + (grad,) if return_grad else out

    def value_and_grad(self, pools, inputs_old, outputs_old, fantasy_rvs=None, num_pending=0, model=None, **kwargs):
        losses, _, grad = self.evaluate(pools, inputs_old, outputs_old, fantasy_rvs=fantasy_rvs, model=model,
                                        num_pending=num_pending, return_grad=True, **kwargs)
        return np.reshape(losses, -1), grad

    def closed_form(self, pools, inputs_old, outputs_old, **kwargs):
        return self.evaluate(pools, inputs_old, outputs_old, subroutine=self.closed_form, **kwargs)

    def monte_carlo(self, pools, inputs_old, outputs_old, fantasy_rvs=None, **kwargs):
        return self.evaluate(pools, inputs_old, outputs_old, fantasy_rvs=fantasy_rvs, subroutine=self.monte_carlo, **kwargs)

    def _monte_carlo(self, means, cov, base_rvs=None, weights=None, incumbents=None, model=None, outputs_old=None,
                     num_fantasies=None, **kwargs):
        """MC estimate from given moments means [..., q, 1] / cov [..., q, q] (the per-class bodies of the
        reference: negative_ei.py:52-74 etc.); jitter_cholesky uses the model's jitter."""
        if model is None:
            model = self.model
        cov = np.asarray(cov, dtype=np.float64)
        lead = cov.shape[:-2]
        q = cov.shape[-1]
        cov3 = cov.reshape(-1, q, q)
        mu2 = np.asarray(means, dtype=np.float64).reshape(-1, q)
        if base_rvs is None:
            base_rvs = self.rng.randn(self.num_fantasies if num_fantasies is None else num_fantasies, q)
        from ... import runtime
        eng = runtime.get_engine(getattr(model, "device", None), self.dtype)
        if getattr(eng, "_fingerprint", None) is None:
            jit = 1e-6 if model is None else model.jitter
            eng.set_model("squared_exponential", np.ones(1), 1.0, None, 0.0, jit)
        prm = self.acq_params(outputs_old=outputs_old, **kwargs)
        losses, _, _, _ = eng.mc_from_moments(prm, mu2, cov3, base_rvs, weights, incumbents)
        return losses.reshape(lead)

    # ---- sampling plumbing --------------------------------------------------------------------------
    def update(self, sess, inputs_old, outputs_old, model=None, **kwargs):
        pass   # only the random-feature variant of the reference does work here (:287-294)

    def prepare(self, sess, inputs_old, outputs_old, parallelism, model=None, num_fantasies=None, dtype=None,
                **kwargs):
        """Base samples for the next evaluation (:297-319): drawn from the self-reseeding host RNG only
        when the pool has more than one point.  Returns (placeholders, feed_dict) keyed by name."""
        if dtype is None:
            dtype = self.dtype
        if num_fantasies is None:
            num_fantasies = self.num_fantasies
        placeholders, feed_dict = dict(), dict()
        if parallelism > 1:
            rvs = self.rng.randn(num_fantasies, parallelism).astype(dtype)
            placeholders["fantasy_rvs"] = "fantasy_rvs"
            feed_dict["fantasy_rvs"] = rvs
            if self.cache_rvs:
                self.fantasy_rvs = rvs
        return placeholders, feed_dict
