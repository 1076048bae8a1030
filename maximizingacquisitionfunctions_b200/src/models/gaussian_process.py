"""Gaussian-process surrogate with the reference's parameterisation and method names
(src/models/gaussian_process.py), evaluated by the CUDA library.

State (``self.state``): ``mean`` (None => empirical mean of the outputs, :229-231),
``log_lenscales`` (scalar or [d]), ``log_amplitude``, ``log_noise`` (None => noiseless) -- plain
NumPy values instead of ``tf.Variable``s.  ``jitter`` defaults to 1e-6 (float64) / 1e-4 (float32)
(:42-44); ``kernel_id`` in {squared_exponential, matern52, matern32} (src/models/kernels.py).
"""
import numpy as np

from ... import runtime
from .abstract_model import abstract_model

_KERNELS = ("squared_exponential", "matern52", "matern32")


class gaussian_process(abstract_model):
    def __init__(self, log_lenscales=np.log(1.0), log_amplitude=np.log(1.0), log_noise=np.log(0.01),
                 mean=0.0, kernel_id="matern52", priors=None, num_states=1, spo_config=None, jitter=None,
                 device=None, **kwargs):
        super().__init__(**kwargs)
        if kernel_id not in _KERNELS:
            raise Exception("Unrecognized kernel function '{:s}'".format(str(kernel_id)))
        if jitter is None:
            jitter = 1e-6 if np.dtype(self.dtype) == np.float64 else 1e-4
        self.jitter = jitter
        self.kernel_id = kernel_id
        self.device = device
        self.priors = dict(priors or {})
        self.spo_config = self.update_dict({
            "method": "L-BFGS-B",
            "var_to_bounds": {
                "log_lenscales": (np.log(1e-2), np.log(2.0)),
                "log_amplitude": (np.log(1e-3), np.log(1e4)),
                "log_noise": (-np.inf, np.log(1.0)),
            },
            "options": {"maxiter": 1000},
        }, spo_config)
        self.states = [{
            "mean": mean,
            "log_lenscales": None if log_lenscales is None else np.asarray(log_lenscales, dtype=np.float64),
            "log_amplitude": float(log_amplitude),
            "log_noise": None if log_noise is None else float(log_noise),
        } for _ in range(max(1, num_states))]
        self.active_state = 0
        self._initial_states = [dict(s) for s in self.states]

    # ---- state ------------------------------------------------------------------------------
    @property
    def state(self):
        return self.states[self.active_state]

    @property
    def num_states(self):
        return len(self.states)

    def reset_state(self):
        self.states[self.active_state] = dict(self._initial_states[self.active_state])

    @property
    def mean(self):
        return self.state.get("mean", None)

    @property
    def log_lenscales(self):
        return self.state["log_lenscales"]

    @property
    def lenscales(self):
        return np.exp(self.log_lenscales)

    @property
    def log_amplitude(self):
        return self.state["log_amplitude"]

    @property
    def amplitude(self):
        return np.exp(self.log_amplitude)

    @property
    def log_noise(self):
        return self.state.get("log_noise", None)

    @property
    def noise(self):
        return None if self.log_noise is None else np.exp(self.log_noise)

    def hyperparameters(self, input_dim):
        """Keyword arguments of Engine.set_model for the active state."""
        ls = np.broadcast_to(np.asarray(self.lenscales, dtype=np.float64).reshape(-1), (input_dim,)) \
            if np.size(self.lenscales) in (1, input_dim) else None
        if ls is None:
            raise ValueError("log_lenscales must be a scalar or have one entry per input dimension")
        return dict(kernel_id=self.kernel_id, lenscales=np.array(ls), amplitude=float(self.amplitude),
                    noise=None if self.noise is None else float(self.noise),
                    mean=None if self.mean is None else float(self.mean), jitter=float(self.jitter))

    # ---- device binding ---------------------------------------------------------------------------
    def bind(self, inputs_old, outputs_old):
        """Engine holding this model's factors for (inputs_old, outputs_old)."""
        X0 = np.asarray(inputs_old, dtype=np.float64)
        Y0 = np.asarray(outputs_old, dtype=np.float64)
        assert X0.ndim == 2 and Y0.shape[0] == X0.shape[0], "observations must be [n, d] / [n, 1]"
        assert Y0.ndim == 1 or Y0.shape[-1] == 1, "Multiple objectives not yet supported"
        eng = runtime.get_engine(self.device, self.dtype)
        return runtime.bind(eng, self.hyperparameters(X0.shape[-1]), X0, Y0.reshape(-1))

    # ---- reference methods ------------------------------------------------------------------------
    def predict(self, inputs_new, inputs_old, outputs_old, full_cov=False, use_rank2=False, chol=None, **kwargs):
        """Posterior (means, cov | var) at `inputs_new` ([m, d] as one set, or [S, q, d] sets).
        Shapes follow _predict_nd (:172-223): means [..., q, 1]; cov [..., q, q] or var [..., q, 1].
        `chol` (the cached factor the reference threads through) is implied by the binding."""
        eng = self.bind(inputs_old, outputs_old)
        X1 = np.asarray(inputs_new, dtype=np.float64)
        rank2 = X1.ndim == 2
        if not full_cov:
            flat = X1.reshape(-1, 1, X1.shape[-1])
            mu, var = eng.posterior(flat, full_cov=False)
            shape = X1.shape[:-1] + (1,)
            return mu.reshape(shape), var.reshape(shape)
        pools = X1[None] if rank2 else X1
        mu, cov = eng.posterior(np.ascontiguousarray(pools), full_cov=True)
        return (mu[0], cov[0]) if rank2 else (mu, cov)

    def compute_cholesky(self, inputs, cov=None, noisy=False, outputs=None, **kwargs):
        """Lower Cholesky factor of K(inputs, inputs) (+ noise) + jitter I (:352-360)."""
        if cov is not None:
            raise NotImplementedError("compute_cholesky(cov=...) is served by the MC kernel (loss._monte_carlo)")
        X0 = np.asarray(inputs, dtype=np.float64)
        y = np.zeros(len(X0)) if outputs is None else np.asarray(outputs, dtype=np.float64).reshape(-1)
        hyper = self.hyperparameters(X0.shape[-1])
        if not noisy:
            hyper["noise"] = None
        eng = runtime.get_engine(self.device, self.dtype)
        runtime.bind(eng, hyper, X0, y)
        return eng.factors()[0]

    def draw_samples(self, num_samples, X1, X0=None, Y0=None, base_rvs=None, **kwargs):
        """Posterior draws [num_samples, m] at one point set X1 [m, d] (:328-349); host RNG supplies
        the base samples unless `base_rvs` [num_samples, m] is given (the reference uses TF's RNG)."""
        if X0 is None or Y0 is None:
            raise NotImplementedError("Sampling from the prior not yet supported")
        X1 = np.asarray(X1, dtype=np.float64).reshape(-1, np.shape(X0)[-1])
        mu, cov = self.predict(X1, X0, Y0, full_cov=True)
        if base_rvs is None:
            base_rvs = self.rng.randn(num_samples, X1.shape[0])
        eng = self.bind(X0, Y0)
        from ...engine import acq_params
        _, _, _, chol = eng.mc_from_moments(acq_params("sr"), mu[None, :, 0], cov[None], np.asarray(base_rvs))
        return mu[:, 0][None, :] + np.asarray(base_rvs) @ chol[0].T

    def fit(self, sess, inputs, outputs, **kwargs):
        raise NotImplementedError(
            "GP MAP fit (gaussian_process.fit, SURVEY 8f-2) is not on the accelerated path yet; "
            "set hyper-parameters explicitly (scripts: --use_true_prior 1)")
