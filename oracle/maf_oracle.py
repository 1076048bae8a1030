"""CPU oracle for the hot path: TEST INFRASTRUCTURE, not product code.

This module restates, op for op, the arithmetic that the reference
(j-wilson/MaximizingAcquisitionFunctions) executes inside TensorFlow 1.x for
multi-start maximisation of Monte-Carlo batch acquisition functions over a GP
posterior.  It exists so that the CUDA path can be checked against it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package never does.

Third-party dependency that holds the reference arithmetic: module
``tensorflow`` 1.x (>=1.5, <2.0; NOT pinned anywhere in the reference tree, NOT
installable here).  The primitives below (cholesky, cholesky_solve, einsum,
reduce_min gradient, Adam kernel, ...) restate TF-1's published semantics on
torch CPU float64 (autograd supplies a gradient that is independent of the
hand-derived CUDA backward).

Pinning status
  * posterior mean / covariance / variance and pdist2: PINNED against the
    reference's own NumPy test oracle (tests/test_models.py:52-87, config
    :229-256, tolerance 100*eps) and scipy ``cdist`` (tests/test_utils.py:43-67)
    -- see tests/golden/make_golden.py and tests/test_oracle.py.
  * q=1 MC UCB -> closed form identity pinned against tests/test_qUCB.py:47-53.
  * acquisition value, posterior moments, Cholesky factor and gradient of
    sum(losses) w.r.t. the pools: PINNED against the reference's OWN code --
    src/losses/{myopic_loss,negative_ei,negative_pi,simple_regret,
    confidence_bound}.py, src/models/{gaussian_process,kernels}.py, src/utils/*,
    src/misc/sharded_fn.py imported unmodified and executed on an eager
    torch-backed stand-in for the ``tensorflow`` module
    (tests/golden/tf_eager_shim, tests/golden/make_golden_acq.py ->
    tests/golden/ref_acq.npz; 13 cases: qEI/qPI soft+hard/qSR/qLCB/qUCB, three
    kernels, weights, incumbents, explicit level, sharding, q=1 closed forms;
    the oracle agrees to 1e-11, tests/test_oracle.py).  What that cannot pin is
    TensorFlow's own kernels (cholesky, cholesky_solve, reduce_min gradient):
    their published TF-1 semantics are what the stand-in implements.
  * Adam trajectory / selected index: no reference test and no reference code
    outside the TF optimizer => "parity unpinned" at that level; anchored on
    the TF-1 ApplyAdam formula cited at adam_step and on np.argmin semantics.

All ``file:line`` citations are relative to the reference root.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch

DT = torch.float64


# --------------------------------------------------------------------------
# host RNG  (src/core/module.py:29-42, 131-144)
# --------------------------------------------------------------------------
class HostRng:
    """Self-reseeding RandomState: every access bumps the seed then reseeds.

    The k-th access (k=1,2,..) of an object constructed with seed ``s0`` is a
    fresh legacy MT19937 ``RandomState(max(1, (s0+k) mod (2**31-1)))``.
    """

    def __init__(self, seed: int):
        self.seed = int(seed)
        self._rng = np.random.RandomState(np.array([seed], dtype=np.int64))

    @property
    def rng(self) -> np.random.RandomState:
        self.seed = max(1, int(np.mod(self.seed + 1, 2 ** 31 - 1)))
        self._rng.seed(self.seed)
        return self._rng


# --------------------------------------------------------------------------
# kernels  (src/models/kernels.py:32-49)
# --------------------------------------------------------------------------
def squared_exponential(r2: torch.Tensor) -> torch.Tensor:
    return torch.exp(-0.5 * r2)


def matern52(r2: torch.Tensor) -> torch.Tensor:
    eps = torch.finfo(r2.dtype).eps
    r2 = 5.0 * torch.clamp(r2, min=eps)          # clip_by_value(r2, eps, inf)
    r = torch.sqrt(r2)
    return (1.0 + r + (1 / 3) * r2) * torch.exp(-r)


def matern32(r2: torch.Tensor) -> torch.Tensor:
    eps = torch.finfo(r2.dtype).eps
    r = torch.sqrt(3.0 * torch.clamp(r2, min=eps))
    return (1.0 + r) * torch.exp(-r)


KERNELS = {
    "squared_exponential": squared_exponential,
    "matern52": matern52,
    "matern32": matern32,
}


# --------------------------------------------------------------------------
# linalg helpers  (src/utils/linalg_ops.py)
# --------------------------------------------------------------------------
def pdist2(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """Direct squared differences (linalg_ops.py:102-109)."""
    diff = A.unsqueeze(-2) - B.unsqueeze(-3)
    return (diff * diff).sum(-1)


def jitter_cholesky(A: torch.Tensor, jitter: float) -> torch.Tensor:
    """chol(A + jitter*I)  (linalg_ops.py:112-150)."""
    eye = torch.eye(A.shape[-1], dtype=A.dtype)
    return torch.linalg.cholesky(A + jitter * eye)


def cholesky_solve_lifted(L: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """Rank-lifting cholesky_solve (linalg_ops.py:184-212), d = rank(B)-rank(L) >= 0.

    B[..., n, q] is flattened to one [n, prod(...)*q] right-hand side, solved
    with two triangular solves (tf.cholesky_solve) and reshaped back.
    """
    if B.dim() == L.dim():
        return torch.cholesky_solve(B, L, upper=False)
    lead = B.shape[:-2]
    n, q = B.shape[-2:]
    Bm = B.reshape(-1, n, q).permute(1, 0, 2).reshape(n, -1)
    Xm = torch.cholesky_solve(Bm, L, upper=False)
    return Xm.reshape(n, -1, q).permute(1, 0, 2).reshape(*lead, n, q)


# --------------------------------------------------------------------------
# GP model  (src/models/gaussian_process.py)
# --------------------------------------------------------------------------
@dataclass
class GP:
    """Hyperparameters of the reference GP in its own parameterisation
    (gaussian_process.py:37-75, 377-411): log-lengthscales (scalar or [d]),
    log-amplitude, log-noise (None => noiseless), constant mean (None =>
    empirical mean of Y0, :229-231), jitter 1e-6 (f64) / 1e-4 (f32)."""

    log_lenscales: object = 0.0
    log_amplitude: float = 0.0
    log_noise: Optional[float] = math.log(0.01)
    mean: Optional[float] = 0.0
    kernel_id: str = "matern52"
    jitter: float = 1e-6

    def lenscales(self) -> torch.Tensor:
        return torch.exp(torch.as_tensor(self.log_lenscales, dtype=DT))

    def amplitude(self) -> torch.Tensor:
        return torch.exp(torch.as_tensor(self.log_amplitude, dtype=DT))

    def noise(self) -> Optional[torch.Tensor]:
        if self.log_noise is None:
            return None
        return torch.exp(torch.as_tensor(self.log_noise, dtype=DT))

    # covar_fn :244-269 (scale by 1/l BEFORE differencing; self-cov of a
    # single argument uses r2 = 0)
    def covar_fn(self, X0: torch.Tensor, X1: Optional[torch.Tensor] = None,
                 noisy: bool = False) -> torch.Tensor:
        kern = KERNELS[self.kernel_id]
        if X1 is not None:
            inv_scales = 1.0 / self.lenscales()
            dist2 = pdist2(inv_scales * X0, inv_scales * X1)
        else:
            dist2 = torch.zeros_like(X0[..., :1])
        cov = (self.amplitude() ** 2) * kern(dist2)
        if not noisy:
            return cov
        return self.noise_fn(cov)

    # noise_fn :272-280
    def noise_fn(self, cov: torch.Tensor) -> torch.Tensor:
        nz = self.noise()
        if nz is None:
            return cov
        if cov.dim() < 2:
            return cov + nz
        return cov + nz * torch.eye(cov.shape[-1], dtype=cov.dtype)

    # mean_fn :226-241
    def mean_value(self, Y0: Optional[torch.Tensor]) -> torch.Tensor:
        if self.mean is None:
            assert Y0 is not None
            return Y0.mean(dim=-2, keepdim=True)
        return torch.as_tensor(self.mean, dtype=DT)

    # compute_cholesky :352-360
    def compute_cholesky(self, X0: torch.Tensor, noisy: bool = True) -> torch.Tensor:
        return jitter_cholesky(self.covar_fn(X0, X0, noisy=noisy), self.jitter)

    # _predict_nd :172-223 (rank-2 X0/Y0; X1 of rank 2 or 3)
    def predict(self, X1: torch.Tensor, X0: torch.Tensor, Y0: torch.Tensor,
                full_cov: bool = False, chol: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        K_10 = self.covar_fn(X1, X0)                       # [..., q, n]
        if chol is None:
            chol = self.compute_cholesky(X0, noisy=True)
        K_01 = K_10.transpose(-1, -2)                      # [..., n, q]
        Beta = cholesky_solve_lifted(chol, K_01).transpose(-1, -2)
        mval = self.mean_value(Y0)
        resid = Y0 - mval                                  # [n, 1]
        Means = mval + Beta @ resid                        # [..., q, 1]
        if full_cov:
            Covs = self.covar_fn(X1, X1) - Beta @ K_01     # NOT symmetrised
            return Means, Covs
        Vars = self.covar_fn(X1) - (Beta * K_10).sum(-1, keepdim=True)
        return Means, Vars


# --------------------------------------------------------------------------
# marginal likelihood + hyper-priors  (gaussian_process.log_likelihood :78-106,
# priors.py:42-73; default priors gaussian_process.py:55-62)
# --------------------------------------------------------------------------
def log_likelihood_data(gp: GP, X0: torch.Tensor, Y0: torch.Tensor) -> torch.Tensor:
    chol = gp.compute_cholesky(X0, noisy=True)
    resid = Y0 - gp.mean_value(Y0)
    gamma = torch.cholesky_solve(resid, chol, upper=False)
    count = X0.shape[0]
    logdet = 2.0 * torch.log(torch.diagonal(chol)).sum()
    return -0.5 * ((resid.transpose(-1, -2) @ gamma).squeeze() + logdet + math.log(2 * math.pi) * count)


def prior_log_normal(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    return -0.5 * torch.square(torch.log(x) / scale) - torch.log(math.sqrt(2 * math.pi) * scale * x)


def prior_horseshoe(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    return torch.log(torch.log(1 + 3 * torch.square(scale / x)))


def negative_log_posterior(gp: GP, X0, Y0) -> torch.Tensor:
    """-(data llh + log_normal(amplitude) + horseshoe(0.1)(noise)): the objective of gaussian_process.fit."""
    X0 = torch.as_tensor(X0, dtype=DT)
    Y0 = torch.as_tensor(Y0, dtype=DT).reshape(-1, 1)
    llh = log_likelihood_data(gp, X0, Y0) + prior_log_normal(gp.amplitude())
    if gp.log_noise is not None:
        llh = llh + prior_horseshoe(gp.noise(), 0.1)
    return -llh


# --------------------------------------------------------------------------
# losses  (src/losses/*.py)
# --------------------------------------------------------------------------
_ISQRT2PI = 1.0 / math.sqrt(2 * math.pi)
_ISQRT2 = 1.0 / math.sqrt(2)


def normal_pdf(x):                       # src/utils/stats_ops.py:39-45
    return _ISQRT2PI * torch.exp(-0.5 * x ** 2)


def normal_cdf(x):                       # src/utils/stats_ops.py:48-51
    return 0.5 * torch.erfc(-_ISQRT2 * x)


def safe_div(num, den):                  # src/utils/math_ops.py:94-109
    both0 = (num == 0) & (den == 0)
    den_safe = torch.where(both0, torch.ones_like(den), den)
    return torch.where(both0, torch.zeros_like(num), num / den_safe)


def soft_less(A, B, temperature):        # src/utils/math_ops.py:56-69
    eps = torch.finfo(A.dtype).eps
    diffs = (B - A) - eps
    inv_temp = torch.reciprocal(torch.as_tensor(temperature, dtype=A.dtype))
    return torch.sigmoid(inv_temp * diffs)


@dataclass
class Acq:
    """Acquisition selector mirroring the four loss classes.

    kind: 'ei' (negative_ei.py), 'pi' (negative_pi.py), 'sr'
    (simple_regret.py), 'cb' (confidence_bound.py).
    """

    kind: str = "ei"
    temperature: Optional[float] = None      # pi only (None => hard indicator)
    beta: float = 2.0                        # cb only
    lower: bool = True                       # cb only
    level: Optional[float] = None            # ei/pi: None => min(outputs_old)


def draw_residuals(chol: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """swap(tensordot(chol[S,q,q], z[M,q], [[-1],[-1]])) -> [S,M,q]
    (myopic_loss.py:247-272)."""
    return torch.tensordot(chol, z, dims=([-1], [-1])).transpose(-1, -2)


def reduce_samples(vals: torch.Tensor, weights: Optional[torch.Tensor]) -> torch.Tensor:
    """(myopic_loss.py:275-284): mean over M, or sum_m w_m v_m."""
    if weights is None:
        return vals.mean(dim=-1)
    return torch.tensordot(weights, vals, dims=([-1], [-1]))


def mc_loss(acq: Acq, means: torch.Tensor, chol: torch.Tensor, z: torch.Tensor,
            outputs_old: torch.Tensor, weights: Optional[torch.Tensor] = None,
            incumbents: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The four ``_monte_carlo`` bodies.  means [S,q,1], chol [S,q,q], z [M,q]
    -> losses [S].

    ei: negative_ei.py:52-74 (level = reduce_min(Y0,-2,keepdims) :31-34)
    pi: negative_pi.py:52-80 (level = reduce_min(Y0) :32-34; soft_less when a
        temperature is set)
    sr: simple_regret.py:32-49
    cb: confidence_bound.py:48-116 (folded-normal residuals; NOTE the
        reference's UCB+incumbents branch calls tf.reduce_max(extrema,
        incumbents) (:75), a bug; the intended maths ``maximum`` is restated).
    """
    resid = draw_residuals(chol, z)                         # [S,M,q]
    mu = means.transpose(-1, -2)                            # [S,1,q]
    if acq.kind == "cb":
        scale = (0.5 * math.pi * acq.beta) ** 0.5
        resid = torch.abs(scale * resid)
        if acq.lower:
            resid = -resid
        samples = resid + mu
        if acq.lower:
            ext = samples.min(dim=-1).values
            if incumbents is not None:
                ext = torch.minimum(ext, incumbents)
        else:
            ext = samples.max(dim=-1).values
            if incumbents is not None:
                ext = torch.maximum(ext, incumbents)
        return reduce_samples(ext, weights)

    samples = resid + mu
    minima = samples.min(dim=-1).values                     # [S,M]
    if incumbents is not None:
        minima = torch.minimum(minima, incumbents)
    if acq.kind == "sr":
        return reduce_samples(minima, weights)
    level = outputs_old.min() if acq.level is None else torch.as_tensor(acq.level, dtype=DT)
    if acq.kind == "ei":
        return -reduce_samples(torch.relu(level - minima), weights)
    if acq.kind == "pi":
        if acq.temperature is None:
            improved = (minima < level).to(minima.dtype)
        else:
            improved = soft_less(minima, level, acq.temperature)
        return -reduce_samples(improved, weights)
    raise ValueError(acq.kind)


def closed_form_loss(acq: Acq, means: torch.Tensor, var: torch.Tensor,
                     outputs_old: torch.Tensor) -> torch.Tensor:
    """The four ``_closed_form`` bodies, means/var [S,1,1] -> [S,1,1].

    ei negative_ei.py:37-49; pi negative_pi.py:38-49; sr simple_regret.py:24-29;
    cb confidence_bound.py:36-45.
    """
    if acq.kind == "sr":
        return means
    if acq.kind == "cb":
        if acq.lower:
            return means - torch.sqrt(acq.beta * var)
        return means + torch.sqrt(acq.beta * var)
    level = outputs_old.min() if acq.level is None else torch.as_tensor(acq.level, dtype=DT)
    resid = level - means
    stddv = torch.sqrt(var).expand_as(resid)
    zvals = safe_div(resid, stddv)
    if acq.kind == "ei":
        return -(stddv * normal_pdf(zvals) + resid * normal_cdf(zvals))
    if acq.kind == "pi":
        return -normal_cdf(zvals)
    raise ValueError(acq.kind)


# --------------------------------------------------------------------------
# the path: pools -> posterior -> chol -> MC loss  (myopic_loss.py:148-179)
# --------------------------------------------------------------------------
@dataclass
class TrainState:
    X0: torch.Tensor
    Y0: torch.Tensor
    chol: torch.Tensor
    gp: GP


def prepare_train(gp: GP, X0, Y0) -> TrainState:
    """bayesian_optimization.appraise_inputs :188-198 -- the training Cholesky
    is computed once and fed back as data (no gradient through it)."""
    X0 = torch.as_tensor(X0, dtype=DT)
    Y0 = torch.as_tensor(Y0, dtype=DT).reshape(-1, 1)
    with torch.no_grad():
        chol = gp.compute_cholesky(X0, noisy=True)
    return TrainState(X0, Y0, chol, gp)


def shard_sizes(S: int, shard_limit: int):
    """sharded_fn with exact_sharding (src/misc/sharded_fn.py:56-83,106-132):
    shard = min(S, limit); S must be a multiple of it."""
    shard = min(S, shard_limit)
    assert S % shard == 0, "exact_sharding requires S to be a multiple of the shard size"
    return [shard] * (S // shard)


def posterior(ts: TrainState, pools: torch.Tensor, full_cov: bool = True):
    return ts.gp.predict(pools, ts.X0, ts.Y0, full_cov=full_cov, chol=ts.chol)


def mc_losses(ts: TrainState, acq: Acq, pools: torch.Tensor, z: torch.Tensor,
              weights=None, incumbents=None, shard_limit: int = 2 ** 12,
              return_extras: bool = False):
    """losses[S] for pools[S,q,d] (q>1 path, myopic_loss.monte_carlo :210-244)."""
    outs, extras = [], []
    start = 0
    for sz in shard_sizes(pools.shape[0], shard_limit):
        P = pools[start:start + sz]
        means, cov = posterior(ts, P, full_cov=True)
        chol = jitter_cholesky(cov, ts.gp.jitter)
        outs.append(mc_loss(acq, means, chol, z, ts.Y0, weights, incumbents))
        if return_extras:
            extras.append((means, cov, chol))
        start += sz
    losses = torch.cat(outs)
    if return_extras:
        return losses, tuple(torch.cat([e[i] for e in extras]) for i in range(3))
    return losses


def closed_form_losses(ts: TrainState, acq: Acq, pools: torch.Tensor,
                       shard_limit: int = 2 ** 16) -> torch.Tensor:
    """losses[S,1,1] for pools[S,1,d] (q==1 path, myopic_loss.closed_form :185-207)."""
    outs = []
    start = 0
    for sz in shard_sizes(pools.shape[0], shard_limit):
        means, var = posterior(ts, pools[start:start + sz], full_cov=False)
        outs.append(closed_form_loss(acq, means, var, ts.Y0))
        start += sz
    return torch.cat(outs)


def fantasy_closed_form_losses(gp: GP, X0, Y, acq: Acq, pools: torch.Tensor) -> torch.Tensor:
    """losses[S] for pools[S,1,d] under fantasy-padded observations (inputs_old [F,1,n,d], outputs_old [F,1,n,1],
    scripts/run_bayesopt_incremental.py:150-179): the q == 1 closed form per fantasy through the rank-4 branch of
    _predict_nd (gaussian_process.py:195-206), then the mean over the fantasy axis (appraise_inputs,
    bayesian_optimization.py:227-229).  Default levels: negative_ei reduces outputs_old along axis -2 (one level per
    fantasy, negative_ei.py:31-34); negative_pi takes tf.reduce_min(outputs_old) over the WHOLE tensor (one level for
    all fantasies, negative_pi.py:32-34).  Pinned by tests/golden/ref_acq.npz (fan_* cases)."""
    Y = np.asarray(Y, dtype=np.float64)
    if acq.kind == "pi" and acq.level is None:
        acq = Acq("pi", temperature=acq.temperature, beta=acq.beta, lower=acq.lower, level=float(Y.min()))
    tot = 0
    for y in Y:
        tot = tot + closed_form_losses(prepare_train(gp, X0, y), acq, pools).reshape(-1)
    return tot / Y.shape[0]


def value_and_grad(ts: TrainState, acq: Acq, X_new, z, X_pend=None,
                   weights=None, incumbents=None, shard_limit: int = 2 ** 12):
    """Loss vector and d(sum of losses)/d(inputs_new).

    pools = concat(tile(pend), new) (bayesian_optimization.prepare :1106-1123);
    TF differentiates a vector loss as its sum (optimizer.minimize :420-426).
    Returns numpy (losses[S], grad[S,q_new,d]).
    """
    Xn = torch.as_tensor(X_new, dtype=DT).clone().requires_grad_(True)
    z = torch.as_tensor(z, dtype=DT)
    if weights is not None:
        weights = torch.as_tensor(weights, dtype=DT)
    if incumbents is not None:
        incumbents = torch.as_tensor(incumbents, dtype=DT)
    S = Xn.shape[0]
    if X_pend is not None and len(X_pend):
        Xp = torch.as_tensor(X_pend, dtype=DT)
        pools = torch.cat([Xp.unsqueeze(0).expand(S, -1, -1), Xn], dim=-2)
    else:
        pools = Xn
    if pools.shape[-2] == 1:
        losses = closed_form_losses(ts, acq, pools).reshape(-1)
    else:
        losses = mc_losses(ts, acq, pools, z, weights, incumbents, shard_limit)
    if not losses.requires_grad:
        # hard-indicator PI is piecewise constant (negative_pi.py:67-69): TF has no gradient
        # to offer (optimizer.minimize raises); report zeros.
        return losses.detach().numpy(), np.zeros(tuple(Xn.shape))
    losses.sum().backward()
    return losses.detach().numpy(), Xn.grad.detach().numpy()


# --------------------------------------------------------------------------
# Adam + clip  (bayesian_optimization._optimize_inputs_sgd :377-482;
# TF-1 ApplyAdam kernel form)
# --------------------------------------------------------------------------
@dataclass
class AdamState:
    m: np.ndarray
    v: np.ndarray
    beta1_power: float
    beta2_power: float
    lr: float = 25e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8


def adam_init(shape, lr=25e-3, beta1=0.9, beta2=0.999, eps=1e-8) -> AdamState:
    return AdamState(np.zeros(shape), np.zeros(shape), beta1, beta2, lr, beta1, beta2, eps)


def adam_step(x: np.ndarray, g: np.ndarray, st: AdamState) -> np.ndarray:
    """One TF-1 Adam update followed by clip_by_value(x, 0, 1) (:428-441).

    alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
    x -= m*alpha/(sqrt(v)+eps); then the beta powers are multiplied.
    """
    alpha = st.lr * math.sqrt(1.0 - st.beta2_power) / (1.0 - st.beta1_power)
    st.m = st.m + (g - st.m) * (1.0 - st.beta1)
    st.v = st.v + (g * g - st.v) * (1.0 - st.beta2)
    x = x - (st.m * alpha) / (np.sqrt(st.v) + st.eps)
    st.beta1_power *= st.beta1
    st.beta2_power *= st.beta2
    return np.clip(x, 0.0, 1.0)


def sgd_run(ts: TrainState, acq: Acq, X_new: np.ndarray, z_steps, X_pend=None, lr=1.0, momentum=0.0):
    """tf.train.GradientDescentOptimizer / MomentumOptimizer(., 0.9) branches of _optimize_inputs_sgd
    (src/agents/bayesian_optimization.py:405-412): learning rate lr * (global_step + 1)^-0.7 with global_step
    counting from 0; Momentum: accum = momentum * accum + g, x -= lr_t * accum; then clip_by_value(x, 0, 1)."""
    x = np.array(X_new, dtype=np.float64)
    accum = np.zeros_like(x)
    trace = []
    for t, z in enumerate(z_steps):
        losses, g = value_and_grad(ts, acq, x, z, X_pend)
        trace.append(losses)
        lr_t = lr * float(t + 1) ** -0.7
        accum = momentum * accum + g
        x = np.clip(x - lr_t * accum, 0.0, 1.0)
    return x, np.asarray(trace)


def adam_run(ts: TrainState, acq: Acq, X_new: np.ndarray, z_steps, X_pend=None,
             lr=25e-3):
    """HOT LOOP B (:462-469): fresh z per step, value+grad, Adam, clip."""
    x = np.array(X_new, dtype=np.float64)
    st = adam_init(x.shape, lr=lr)
    trace = []
    for z in z_steps:
        losses, g = value_and_grad(ts, acq, x, z, X_pend)
        trace.append(losses)
        x = adam_step(x, g, st)
    return x, np.asarray(trace)


# --------------------------------------------------------------------------
# synthetic configs  (SURVEY.md 8d / BASELINE.md 3)
# --------------------------------------------------------------------------
def synthetic_problem(n, d, q, S, M, seed=0, kernel_id="matern52", task=None):
    """Seeded inputs shared by the tests, bench and the CUDA path:
    X0=RS(seed).rand(n,d); y0=RS(seed+1).randn(n,1) (or task values + noise);
    X=RS(seed+2).rand(S,q,d); z=RS(seed+3).randn(M,q); l=sqrt(d)/4, a=1,
    noise=1e-3, mean=0, jitter=1e-6 (scripts/run_bayesopt.py:71-74,102-103)."""
    RS = np.random.RandomState
    X0 = RS(seed).rand(n, d)
    if task is None:
        y0 = RS(seed + 1).randn(n, 1)
    else:
        y0 = np.asarray(task(X0)).reshape(n, 1) + math.sqrt(1e-3) * RS(seed + 1).randn(n, 1)
    X = RS(seed + 2).rand(S, q, d)
    z = RS(seed + 3).randn(M, q)
    gp = GP(log_lenscales=math.log(math.sqrt(d) / 4), log_amplitude=0.0,
            log_noise=math.log(1e-3), mean=0.0, kernel_id=kernel_id, jitter=1e-6)
    return gp, X0, y0, X, z
