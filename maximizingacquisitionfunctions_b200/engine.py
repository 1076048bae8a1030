"""Host-side handle on one GPU context of the CUDA library (NumPy in / NumPy out).

This is the thin layer the ``src``-compatible mirror (``.src.models`` / ``.src.losses`` /
``.src.agents``) sits on.  Every number comes from the kernels behind the C ABI
(include/maf.h); nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from . import _native as nat


class MafError(RuntimeError):
    pass


class DeviceBuffer:
    """A raw device allocation owned by an Engine (used by the bench for HBM-resident inputs)."""

    def __init__(self, engine: "Engine", nbytes: int):
        self.engine = engine
        self.nbytes = int(nbytes)
        self.ptr = engine._lib.maf_dev_alloc(engine._ctx, self.nbytes)
        if not self.ptr:
            raise MafError(engine._err())

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.engine._ck(self.engine._lib.maf_h2d(self.engine._ctx, self.ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        return self

    def download(self, shape, dtype=np.float64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        self.engine._ck(self.engine._lib.maf_d2h(self.engine._ctx, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.engine._lib.maf_dev_free(self.engine._ctx, self.ptr)
            self.ptr = None


def acq_params(kind: str, level: Optional[float] = None, temperature: Optional[float] = None,
               beta: float = 2.0, lower: bool = True) -> nat.AcqParams:
    """kind in {'ei','pi','sr','cb'}; level None => min(y0); temperature None => hard PI."""
    return nat.AcqParams(nat.ACQ_IDS[kind], 1 if lower else 0,
                         float("nan") if level is None else float(level),
                         float("nan") if temperature is None else float(temperature), float(beta))


class Engine:
    def __init__(self, device: int = 0, dtype: str = "float64"):
        self._lib = nat.load()
        self._ctx = C.c_void_p()
        code = {"float64": nat.F64, "float32": nat.F32}[str(dtype)]
        rc = self._lib.maf_create(int(device), code, C.byref(self._ctx))
        if rc != 0:
            msg = self._err()
            if self._ctx:
                self._lib.maf_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise MafError("maf_create failed: " + msg)
        self.device = int(device)
        self.dtype = str(dtype)
        self._pinned = []
        self.d = None
        self.n = None
        self._fingerprint = None
        self._fantasy_key = None
        self.last_eval_shape = None        # (S, q, d) of the last value_and_grad: what best_record_dev reports on

    # -- plumbing ---------------------------------------------------------------------------
    def _err(self) -> str:
        return (self._lib.maf_last_error(self._ctx) or b"").decode()

    def _ck(self, rc: int):
        if rc != 0:
            raise MafError(self._err())

    def close(self):
        if getattr(self, "_ctx", None):
            for p in getattr(self, "_pinned", []):
                self._lib.maf_host_free(self._ctx, p)
            self._pinned = []
            self._lib.maf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self._lib.maf_sync(self._ctx))

    @property
    def version(self) -> str:
        return self._lib.maf_version().decode()

    @property
    def launch_count(self) -> int:
        return int(self._lib.maf_launch_count(self._ctx))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """NumPy array backed by page-locked host memory (freed with the engine)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.maf_host_alloc(self._ctx, nbytes)
        if not p:
            raise MafError(self._err())
        self._pinned.append(p)
        buf = (C.c_char * max(nbytes, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def timer_start(self):
        self._ck(self._lib.maf_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        self._ck(self._lib.maf_timer_stop(self._ctx, C.byref(ms)))
        return float(ms.value)

    # -- model / training set ------------------------------------------------------------------
    def set_model(self, kernel_id: str, lenscales: Sequence[float], amplitude: float,
                  noise: Optional[float], mean: Optional[float], jitter: float):
        ls = nat.as_f64(np.atleast_1d(lenscales))
        self.d = int(ls.shape[0])
        self._fingerprint = self._fantasy_key = None      # the factors (and fantasies) belong to the old model
        self._ck(self._lib.maf_set_model(
            self._ctx, nat.KERNEL_IDS[kernel_id], self.d, nat.ptr(ls), float(amplitude),
            float("nan") if noise is None else float(noise),
            float("nan") if mean is None else float(mean), float(jitter)))

    def set_train(self, X0: np.ndarray, y0: np.ndarray):
        X0 = nat.as_f64(X0)
        y0 = nat.as_f64(np.reshape(y0, -1))
        assert X0.ndim == 2 and X0.shape[1] == self.d and y0.shape[0] == X0.shape[0]
        self.n = int(X0.shape[0])
        self._fantasy_key = None                          # maf_set_train resets F to 0 on the device
        self._ck(self._lib.maf_set_train(self._ctx, self.n, nat.ptr(X0), nat.ptr(y0)))

    def factors(self):
        n = self.n
        L0, Linv, gamma = np.empty((n, n)), np.empty((n, n)), np.empty(n)
        self._ck(self._lib.maf_get_factors(self._ctx, nat.ptr(L0), nat.ptr(Linv), nat.ptr(gamma)))
        return L0, Linv, gamma

    def gp_nll(self):
        """(data negative log-likelihood, gradient w.r.t. [log lengthscales(d), log amplitude, log noise, mean])."""
        nll = C.c_double(0.0)
        grad = np.zeros(self.d + 3)
        self._ck(self._lib.maf_gp_nll(self._ctx, C.byref(nll), nat.ptr(grad)))
        return float(nll.value), grad

    # -- posterior ---------------------------------------------------------------------------------
    def posterior(self, X: np.ndarray, full_cov: bool = True):
        X = nat.as_f64(X)
        S, q, d = X.shape
        assert d == self.d
        mu = np.empty((S, q, 1))
        cov = np.empty((S, q, q) if full_cov else (S, q, 1))
        self._ck(self._lib.maf_posterior(self._ctx, S, q, nat.ptr(X), 1 if full_cov else 0, nat.ptr(mu), nat.ptr(cov)))
        return mu, cov

    # -- acquisition -------------------------------------------------------------------------------
    def value_and_grad(self, prm: nat.AcqParams, X: np.ndarray, z: Optional[np.ndarray] = None,
                       p_pending: int = 0, weights=None, incumbents=None, need_grad: bool = True,
                       out_loss: Optional[np.ndarray] = None, out_grad: Optional[np.ndarray] = None):
        """X[S][q][d] pools (pending rows first); returns (losses[S], grad[S][q-p][d] | None)."""
        X = nat.as_f64(X)
        S, q, d = X.shape
        assert d == self.d
        M = 0
        if z is not None:
            z = nat.as_f64(z)
            assert z.ndim == 2 and z.shape[1] == q
            M = int(z.shape[0])
        if weights is not None:
            weights = nat.as_f64(weights)
            assert weights.shape == (M,)
        if incumbents is not None:
            incumbents = nat.as_f64(incumbents)
            assert incumbents.shape == (M,)
        loss = np.empty(S) if out_loss is None else out_loss
        grad = (np.empty((S, q - p_pending, d)) if out_grad is None else out_grad) if need_grad else None
        assert loss.shape == (S,) and (grad is None or grad.shape == (S, q - p_pending, d))
        self._ck(self._lib.maf_acq_value_grad(self._ctx, C.byref(prm), S, q, int(p_pending), nat.ptr(X), M,
                                              nat.ptr(z), nat.ptr(weights), nat.ptr(incumbents),
                                              nat.ptr(loss), nat.ptr(grad)))
        self.last_eval_shape = (S, q, d)
        return loss, grad

    def value_and_grad_dev(self, prm: nat.AcqParams, S: int, q: int, p_pending: int, dX: DeviceBuffer, M: int,
                           dz: Optional[DeviceBuffer], dloss: DeviceBuffer, dgrad: Optional[DeviceBuffer]):
        """Asynchronous, device-resident variant (no copies, no sync)."""
        self._ck(self._lib.maf_acq_value_grad_dev(self._ctx, C.byref(prm), S, q, int(p_pending), dX.ptr, M,
                                                  dz.ptr if dz else None, None, None, dloss.ptr,
                                                  dgrad.ptr if dgrad else None))
        self.last_eval_shape = (S, q, self.d)

    def mc_from_moments(self, prm: nat.AcqParams, mu: np.ndarray, cov: np.ndarray, z: np.ndarray,
                        weights=None, incumbents=None, need_grad: bool = False):
        """MC estimator from given moments mu[S,q(,1)], cov[S,q,q]; returns (loss, mubar, covbar, chol)."""
        cov = nat.as_f64(cov)
        S, q, _ = cov.shape
        mu = nat.as_f64(np.reshape(mu, (S, q)))
        z = nat.as_f64(z)
        M = int(z.shape[0])
        assert z.shape == (M, q)
        weights = None if weights is None else nat.as_f64(weights)
        incumbents = None if incumbents is None else nat.as_f64(incumbents)
        loss, chol = np.empty(S), np.empty((S, q, q))
        mubar = np.empty((S, q)) if need_grad else None
        covbar = np.empty((S, q, q)) if need_grad else None
        self._ck(self._lib.maf_mc_from_moments(self._ctx, C.byref(prm), S, q, nat.ptr(mu), nat.ptr(cov), M,
                                               nat.ptr(z), nat.ptr(weights), nat.ptr(incumbents), nat.ptr(loss),
                                               nat.ptr(mubar), nat.ptr(covbar), nat.ptr(chol)))
        return loss, mubar, covbar, chol

    # -- fantasy-conditioned closed form (rank-4 observation sets) -----------------------------------
    def set_fantasies(self, Y: np.ndarray):
        """Y[F][n]: F output vectors over the loaded training inputs."""
        Y = nat.as_f64(Y)
        assert Y.ndim == 2 and Y.shape[1] == self.n
        self.F = int(Y.shape[0])
        self._ck(self._lib.maf_set_fantasies(self._ctx, self.F, nat.ptr(Y)))

    def fantasies_value_and_grad(self, prm: nat.AcqParams, X: np.ndarray, need_grad: bool = True):
        X = nat.as_f64(np.reshape(X, (-1, self.d)))
        S = X.shape[0]
        loss = np.empty(S)
        grad = np.empty((S, self.d)) if need_grad else None
        self._ck(self._lib.maf_fantasies_value_grad(self._ctx, C.byref(prm), S, nat.ptr(X), nat.ptr(loss), nat.ptr(grad)))
        return loss, grad

    def fantasies_posterior(self, X: np.ndarray):
        X = nat.as_f64(np.reshape(X, (-1, self.d)))
        S = X.shape[0]
        mu, var = np.empty((self.F, S)), np.empty(S)
        self._ck(self._lib.maf_fantasies_posterior(self._ctx, S, nat.ptr(X), nat.ptr(mu), nat.ptr(var)))
        return mu, var

    def last_extras(self, S: int, q: int, want_chol: bool = True):
        mu, cov = np.empty((S, q, 1)), np.empty((S, q, q))
        chol = np.empty((S, q, q)) if want_chol else None
        self._ck(self._lib.maf_last_extras(self._ctx, nat.ptr(mu), nat.ptr(cov), nat.ptr(chol)))
        return mu, cov, chol

    def argmin_last(self):
        idx, val = C.c_int64(-1), C.c_double(0.0)
        self._ck(self._lib.maf_argmin_last(self._ctx, C.byref(idx), C.byref(val)))
        return int(idx.value), float(val.value)

    def best_record_dev(self, index_offset: int, d_rec_ptr: int):
        """[loss*, index_offset + argmin, X*[q][d]] of the last evaluation into device memory (asynchronous)."""
        self._ck(self._lib.maf_best_record_dev(self._ctx, int(index_offset), C.c_void_p(int(d_rec_ptr))))

    # -- Adam loop -----------------------------------------------------------------------------------
    OPTIMIZERS = {"adam": 0, "gd": 1, "momentum": 2}

    def adam_begin(self, prm: nat.AcqParams, X: np.ndarray, p_pending: int = 0, lr: float = 25e-3,
                   beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, method: str = "adam",
                   momentum: float = 0.9):
        """method: 'adam' (default), 'gd' or 'momentum' (lr * (t+1)^-0.7 decay, bayesian_optimization.py:405-412)."""
        X = nat.as_f64(X)
        S, q, d = X.shape
        self._adam_shape = (S, q, d)
        self._ck(self._lib.maf_adam_begin(self._ctx, C.byref(prm), S, q, int(p_pending), nat.ptr(X),
                                          float(lr), float(beta1), float(beta2), float(eps)))
        if method != "adam":
            self._ck(self._lib.maf_adam_set_method(self._ctx, self.OPTIMIZERS[method], float(momentum)))

    def adam_steps(self, z_steps, want_trace: bool = False, steps: Optional[int] = None):
        """z_steps[steps][M][q]; for q == 1 pass z_steps=None and `steps` (closed-form loss)."""
        if z_steps is None:
            assert self._adam_shape[1] == 1 and steps is not None
            M = 0
        else:
            z_steps = nat.as_f64(z_steps)
            steps, M, q = z_steps.shape
            assert q == self._adam_shape[1]
        trace = np.empty((steps, self._adam_shape[0])) if want_trace else None
        self._ck(self._lib.maf_adam_steps(self._ctx, steps, M, nat.ptr(z_steps), nat.ptr(trace)))
        return trace

    # -- batched L-BFGS loop ---------------------------------------------------------------------------
    def lbfgs_run(self, prm: nat.AcqParams, X: np.ndarray, z: Optional[np.ndarray] = None, p_pending: int = 0,
                  fantasies: bool = False, maxiter: int = 1000, ftol: float = 1e6 * np.finfo("float64").eps,
                  pgtol: float = 1e-5, time_limit: float = float("inf"), want_trace: bool = True):
        """Per-restart box-constrained L-BFGS on the device (maf_lbfgs_run).  X[S][q][d] starts (pending rows first);
        returns (X_best[S][q][d], losses[S], trace[evals][S] | None, info dict)."""
        X = np.array(nat.as_f64(X))                      # in/out buffer
        S, q, d = X.shape
        assert d == self.d
        M = 0
        if z is not None:
            z = nat.as_f64(z)
            assert z.ndim == 2 and z.shape[1] == q
            M = int(z.shape[0])
        loss = np.empty(S)
        info = np.zeros(2, dtype=np.int32)
        trace = np.empty((int(maxiter), S)) if want_trace else None
        tl = float("nan") if (time_limit is None or not np.isfinite(time_limit)) else float(time_limit)
        self._ck(self._lib.maf_lbfgs_run(self._ctx, C.byref(prm), S, q, int(p_pending), nat.ptr(X), M, nat.ptr(z),
                                         1 if fantasies else 0, int(maxiter), float(ftol), float(pgtol), tl, nat.ptr(loss),
                                         info.ctypes.data_as(C.c_void_p), nat.ptr(trace)))
        self.last_eval_shape = None
        return X, loss, (None if trace is None else trace[:int(info[0])]), {"evaluations": int(info[0]), "not_converged": int(info[1])}

    def adam_peek(self) -> np.ndarray:
        X = np.empty(self._adam_shape)
        self._ck(self._lib.maf_adam_peek(self._ctx, nat.ptr(X)))
        return X

    def adam_end(self) -> np.ndarray:
        X = np.empty(self._adam_shape)
        self._ck(self._lib.maf_adam_end(self._ctx, nat.ptr(X)))
        return X

    # -- profiling -------------------------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._ck(self._lib.maf_profile_enable(self._ctx, 1 if on else 0))

    def profile_reset(self):
        self._ck(self._lib.maf_profile_reset(self._ctx))

    def profile_read(self):
        ms = np.zeros(len(nat.PROF_SLOTS))
        ln = np.zeros(len(nat.PROF_SLOTS), dtype=np.int64)
        self._ck(self._lib.maf_profile_read(self._ctx, nat.ptr(ms), ln.ctypes.data_as(C.c_void_p)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(nat.PROF_SLOTS)}

    def set_gemm_mode(self, mode: int):
        """0 = FP64 DMMA contractions, 1 = INT8 digit-sliced contractions (effective at the next set_train)."""
        self._ck(self._lib.maf_set_gemm_mode(self._ctx, int(mode)))
        self._fingerprint = None

    def set_digits(self, ns_fwd: int, ns_bwd: int):
        """Digit planes of the forward / reverse INT8 contraction (7 / 7: FP64-grade, 28 + 28 products)."""
        self._ck(self._lib.maf_set_digits(self._ctx, int(ns_fwd), int(ns_bwd)))

    def set_level_switch(self, on: bool, tol: Optional[float] = None):
        """Error-bounded level switch of the contractions (see include/maf.h); effective at the next set_train."""
        self._ck(self._lib.maf_set_level_switch(self._ctx, 1 if on else 0, float("nan") if tol is None else float(tol)))
        self._fingerprint = None

    def level_switch_stats(self, reset: bool = False):
        counts = np.zeros(4, dtype=np.int64)
        errs = np.zeros(3)
        self._ck(self._lib.maf_level_switch_stats(self._ctx, counts.ctypes.data_as(C.c_void_p), nat.ptr(errs), 1 if reset else 0))
        return {"tested_forward": int(counts[0]), "recomputed_forward": int(counts[1]), "tested_reverse": int(counts[2]),
                "recomputed_reverse": int(counts[3]), "restarts": int(max(counts[0], counts[2])),
                "err_cov_abs": float(errs[0]), "err_grad_per_scale": float(errs[1]), "tol": float(errs[2])}

    def gemm_nt_i8(self, A: np.ndarray, T: np.ndarray, tri: int = 0, ns: int = 7, lmax: int = 8) -> np.ndarray:
        A, T = nat.as_f64(A), nat.as_f64(T)
        J, K = A.shape
        N, K2 = T.shape
        assert K == K2
        dA, dT = self.alloc(A.nbytes).upload(A), self.alloc(T.nbytes).upload(T)
        dC = self.alloc(J * N * 8)
        self._ck(self._lib.maf_gemm_nt_i8_dev(self._ctx, J, N, K, dA.ptr, K, dT.ptr, K, dC.ptr, N, int(tri), ns, lmax))
        self.sync()
        out = dC.download((J, N))
        for b in (dA, dT, dC):
            b.free()
        return out

    # -- raw GEMM (tests / micro-benchmarks) --------------------------------------------------------------
    def gemm_nt(self, A: np.ndarray, T: np.ndarray, tri: int = 0) -> np.ndarray:
        A, T = nat.as_f64(A), nat.as_f64(T)
        J, K = A.shape
        N, K2 = T.shape
        assert K == K2
        dA, dT = self.alloc(A.nbytes).upload(A), self.alloc(T.nbytes).upload(T)
        dC = self.alloc(J * N * 8)
        self._ck(self._lib.maf_gemm_nt_dev(self._ctx, J, N, K, dA.ptr, K, dT.ptr, K, dC.ptr, N, int(tri)))
        self.sync()
        out = dC.download((J, N))
        for b in (dA, dT, dC):
            b.free()
        return out
