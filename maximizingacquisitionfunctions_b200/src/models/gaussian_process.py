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
        if priors:
            # the MAP objective below hard-codes the reference's defaults (log-normal amplitude, horseshoe(0.1) noise,
            # :55-62); silently ignoring user priors would change what `fit` optimises
            raise NotImplementedError("custom hyper-priors are outside the accelerated path; the defaults of "
                                      "gaussian_process.py:55-62 are built in")
        self.priors = dict()
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
        """Engine holding this model's factors for (inputs_old, outputs_old).  Rank-4 ("fantasy-padded")
        observation sets [F,1,n,d] / [F,1,n,1] share their inputs (the reference solves with chol[0,0],
        :195-200): the factor is built once and the F output vectors are loaded as fantasies."""
        X0 = np.asarray(inputs_old, dtype=np.float64)
        Y0 = np.asarray(outputs_old, dtype=np.float64)
        if X0.ndim == 4:
            assert X0.shape[1] == 1 and Y0.shape[:3] == X0.shape[:3] and Y0.shape[-1] == 1
            if not np.array_equal(X0, np.broadcast_to(X0[:1], X0.shape)):
                raise ValueError("fantasy-padded observation sets must share their inputs")
            eng = self.bind(X0[0, 0], Y0[0, 0])
            Yf = np.ascontiguousarray(Y0[:, 0, :, 0])
            key = (eng._fingerprint, Yf.tobytes())
            if getattr(eng, "_fantasy_key", None) != key:
                eng.set_fantasies(Yf)
                eng._fantasy_key = key
            return eng
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
        if np.ndim(inputs_old) == 4:                  # [F, S, q=1, 1] means; the variance is shared (:195-223)
            assert not full_cov or X1.reshape(-1, X1.shape[-1]).shape[0] == 1 or X1.shape[-2] == 1
            mu, var = eng.fantasies_posterior(X1.reshape(-1, X1.shape[-1]))
            F, S = mu.shape
            lead = (F,) + (X1.shape[:-1] if X1.ndim == 3 else (S,)) + (1,)
            return mu.reshape(lead), np.broadcast_to(var.reshape(lead[1:]), lead).copy()
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
        if np.ndim(X0) == 4:
            # one shared base sample per draw across the F fantasies, as the reference's broadcasted
            # tensordot does (:340-349): samples [F, 1, num_samples... ] -> [F, 1, 1, num_samples]
            assert X1.shape[0] == 1, "fantasy-conditioned draws are taken at one point"
            mu, var = self.predict(X1[None], X0, Y0)             # [F,1,1,1]
            if base_rvs is None:
                base_rvs = self.rng.randn(num_samples, 1)
            sd = np.sqrt(var + self.jitter)
            return mu + sd * np.asarray(base_rvs).reshape(1, 1, 1, -1)
        mu, cov = self.predict(X1, X0, Y0, full_cov=True)
        if base_rvs is None:
            base_rvs = self.rng.randn(num_samples, X1.shape[0])
        eng = self.bind(X0, Y0)
        from ...engine import acq_params
        _, _, _, chol = eng.mc_from_moments(acq_params("sr"), mu[None, :, 0], cov[None], np.asarray(base_rvs))
        return mu[:, 0][None, :] + np.asarray(base_rvs) @ chol[0].T

    # ---- MAP fit (:283-325) -------------------------------------------------------------------------
    def _pack(self, d):
        """Free parameters of the active state as (names, x0, bounds): every state entry that is not None."""
        st = self.state
        bounds_cfg = self.spo_config.get("var_to_bounds", {})
        names, x0, bounds = [], [], []
        if st.get("mean", None) is not None:
            names.append(("mean", 1))
            x0.append([float(st["mean"])])
            bounds.append([(None, None)])
        for key in ("log_lenscales", "log_amplitude", "log_noise"):
            if st.get(key, None) is None:
                continue
            v = np.atleast_1d(np.asarray(st[key], dtype=np.float64)).ravel()
            lo, hi = bounds_cfg.get(key, (None, None))
            lo = None if lo is None or np.isneginf(lo) else float(lo)
            hi = None if hi is None or np.isposinf(hi) else float(hi)
            names.append((key, v.size))
            x0.append(v)
            bounds.append([(lo, hi)] * v.size)
        return names, np.concatenate(x0), [b for bb in bounds for b in bb]

    def _unpack(self, names, x):
        st, k = self.state, 0
        for key, size in names:
            v = np.array(x[k:k + size])
            k += size
            if key == "log_lenscales":
                st[key] = v.reshape(np.shape(st[key])) if np.ndim(st[key]) else np.asarray(v[0])
            else:
                st[key] = float(v[0])

    def negative_log_posterior(self, inputs, outputs):
        """-(data log-likelihood + hyper-priors) and its gradient w.r.t. the packed free parameters.
        Data term and gradient from the device (maf_gp_nll); default priors (:55-62): log-normal on the
        amplitude, horseshoe(0.1) on the noise variance (src/models/priors.py:42-73)."""
        X0 = np.asarray(inputs, dtype=np.float64)
        d = X0.shape[-1]
        eng = self.bind(X0, outputs)
        nll, g = eng.gp_nll()                         # [dlog_ls(d), dlog_amp, dlog_noise, dmean]
        la = float(self.log_amplitude)
        nll -= -0.5 * la * la - la - 0.5 * np.log(2 * np.pi)          # log_normal(scale=1)(exp(la))
        g_amp = g[d] + (la + 1.0)
        g_noise = g[d + 1]
        if self.log_noise is not None:
            # horseshoe(0.1)(x = exp(ln)) = log(log(1 + u)), u = 3 (0.1/x)^2, in the overflow-safe form
            # log1p(u) = logaddexp(0, t), u/(1+u) = sigmoid(t), t = log u  (the prior has no finite mode:
            # it keeps rewarding smaller noise, so L-BFGS-B can walk far down the unbounded side)
            ln = float(self.log_noise)
            t = np.log(3.0 * 0.1 ** 2) - 2.0 * ln
            l1p = np.logaddexp(0.0, t)
            nll -= np.log(l1p)
            g_noise = g_noise + 2.0 * (1.0 / l1p) * (1.0 / (1.0 + np.exp(-t)))
        grads = {"mean": np.array([g[d + 2]]), "log_amplitude": np.array([g_amp]), "log_noise": np.array([g_noise]),
                 "log_lenscales": g[:d] if np.size(self.log_lenscales) == d and np.ndim(self.log_lenscales) else
                 np.array([g[:d].sum()])}
        return float(nll), grads

    def fit(self, sess, inputs, outputs, var_list=None, spo_config=None, feed_dict=None, **kwargs):
        """MAP estimate of the active state's free parameters by L-BFGS-B with the reference's box
        bounds (:65-75); SciPy drives, the device evaluates.  Returns the fitted values."""
        from scipy.optimize import minimize
        inputs = np.asarray(inputs, dtype=np.float64)
        outputs = np.asarray(outputs, dtype=np.float64)
        assert inputs.ndim == outputs.ndim == 2, "Tensor rank should be 2"
        cfg = self.update_dict(self.spo_config, spo_config, as_copy=True)
        names, x0, bounds = self._pack(inputs.shape[-1])
        x0 = np.array([min(max(v, -np.inf if lo is None else lo), np.inf if hi is None else hi)
                       for v, (lo, hi) in zip(x0, bounds)])

        def fun(x):
            self._unpack(names, x)
            val, grads = self.negative_log_posterior(inputs, outputs)
            return val, np.concatenate([np.ravel(grads[k])[:size] for k, size in names])

        res = minimize(fun, x0, jac=True, method=cfg.get("method", "L-BFGS-B"), bounds=bounds,
                       options=dict(cfg.get("options", {})))
        self._unpack(names, res.x)
        return [self.state[k] for k, _ in names]
