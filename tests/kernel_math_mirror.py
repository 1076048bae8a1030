"""NumPy mirror of the arithmetic the CUDA kernels perform (same formulas, same order of
stages, no autograd).  Test-only: it lets the hand-derived reverse pass be checked against
the oracle's autograd on a CPU box, before the kernels ever run.

Stages (one function per kernel):
  cross_fwd      csrc/cov_kernels.cuh  cross_fwd_kernel
  gemm           csrc/gemm_f64.cuh     gemm_nt_f64_kernel (V^T = K10 Linv^T ; Kbar = Vbar^T Linv)
  posterior      csrc/posterior.cuh    posterior_kernel
  mc             csrc/mc_kernels.cuh   mc_acq_kernel
  vbar           csrc/posterior.cuh    vbar_kernel
  cross_bwd      csrc/cov_kernels.cuh  cross_bwd_kernel
"""
import math

import numpy as np
import scipy.linalg as spla

EPS = np.finfo(np.float64).eps


def kern_val(kid, r2):
    if kid == "squared_exponential":
        return np.exp(-0.5 * r2)
    if kid == "matern52":
        u = 5.0 * np.maximum(r2, EPS)
        r = np.sqrt(u)
        return (1.0 + r + (1.0 / 3.0) * u) * np.exp(-r)
    r = np.sqrt(3.0 * np.maximum(r2, EPS))
    return (1.0 + r) * np.exp(-r)


def kern_dr2(kid, r2):
    if kid == "squared_exponential":
        return -0.5 * np.exp(-0.5 * r2)
    if kid == "matern52":
        r = np.sqrt(5.0 * r2)
        out = -(5.0 / 6.0) * (1.0 + r) * np.exp(-r)
    else:
        r = np.sqrt(3.0 * r2)
        out = -1.5 * np.exp(-r)
    return np.where(r2 < EPS, 0.0, out)


class Mirror:
    def __init__(self, kernel_id, lenscales, amplitude, noise, mean, jitter, X0, y0):
        self.kid = kernel_id
        self.d = X0.shape[1]
        self.il = 1.0 / (np.ones(self.d) * np.asarray(lenscales, dtype=np.float64))
        self.amp2 = amplitude ** 2
        self.jitter = jitter
        self.X0s = X0 * self.il
        n = len(X0)
        r2 = ((self.X0s[:, None, :] - self.X0s[None, :, :]) ** 2).sum(-1)
        K00 = self.amp2 * kern_val(self.kid, r2)
        K00[np.diag_indices(n)] = (np.diag(K00) + (noise or 0.0)) + jitter
        self.L0 = spla.cholesky(K00, lower=True)
        self.Linv = spla.solve_triangular(self.L0, np.eye(n), lower=True)
        self.mean = float(np.mean(y0)) if mean is None else mean
        self.gamma = self.Linv @ (y0.reshape(-1) - self.mean)
        self.ymin = float(y0.min())

    # forward ------------------------------------------------------------------------------------------
    def forward(self, X):
        S, q, d = X.shape
        Xs = X * self.il
        r2 = ((Xs[:, :, None, :] - self.X0s[None, None, :, :]) ** 2).sum(-1)      # [S,q,n]
        K10 = self.amp2 * kern_val(self.kid, r2)
        Vt = K10 @ self.Linv.T                                                    # [S,q,n]
        mu = self.mean + Vt @ self.gamma                                          # [S,q]
        r2s = ((Xs[:, :, None, :] - Xs[:, None, :, :]) ** 2).sum(-1)
        idx = np.arange(q)
        r2s[:, idx, idx] = 0.0
        Sig = self.amp2 * kern_val(self.kid, r2s) - Vt @ Vt.transpose(0, 2, 1)
        return Xs, r2, Vt, mu, Sig, r2s

    def mc(self, acq, mu, Sig, z, weights=None, incumbents=None, level=None, temperature=None,
           beta=2.0, lower=True):
        """Returns loss[S], mubar[S,q], Sigbar[S,q,q], chol[S,q,q]."""
        S, q = mu.shape
        M = z.shape[0]
        loss = np.zeros(S)
        mubar = np.zeros((S, q))
        Sigbar = np.zeros((S, q, q))
        chols = np.zeros((S, q, q))
        level = self.ymin if level is None else level
        cbs = (0.5 * math.pi * beta) ** 0.5
        use_max = acq == "cb" and not lower
        w = np.full(M, 1.0 / M) if weights is None else weights
        for s in range(S):
            A = np.tril(Sig[s]) + self.jitter * np.eye(q)
            L = np.zeros((q, q))
            A = A.copy()
            for k in range(q):                         # right-looking, as in the kernel
                lkk = math.sqrt(A[k, k])
                L[k, k] = lkk
                L[k + 1:, k] = A[k + 1:, k] / lkk
                for j in range(k + 1, q):
                    A[j:, j] -= L[j:, k] * L[j, k]
            chols[s] = L
            e = z @ L.T                                # [M,q]
            if acq == "cb":
                t = cbs * e
                y = mu[s] - np.abs(t) if lower else mu[s] + np.abs(t)
                sg = np.sign(t) * (-cbs if lower else cbs)
            else:
                y = mu[s] + e
                sg = np.ones_like(e)
            ext = y.max(1) if use_max else y.min(1)
            mask = (y == ext[:, None])
            cnt = mask.sum(1)
            passed = np.ones(M, dtype=bool)
            ext2 = ext
            if incumbents is not None:
                if use_max:
                    passed = ext >= incumbents
                    ext2 = np.maximum(ext, incumbents)
                else:
                    passed = ext <= incumbents
                    ext2 = np.minimum(ext, incumbents)
            if acq == "ei":
                t = level - ext2
                contrib = -w * np.maximum(t, 0.0)
                dl = np.where(t > 0, w, 0.0)
            elif acq == "pi":
                if temperature is None:
                    contrib = np.where(ext2 < level, -w, 0.0)
                    dl = np.zeros(M)
                else:
                    a = (1.0 / temperature) * ((level - ext2) - EPS)
                    sig = 1.0 / (1.0 + np.exp(-a))
                    contrib = -w * sig
                    dl = w * sig * (1.0 - sig) * (1.0 / temperature)
            else:
                contrib = w * ext2
                dl = w
            loss[s] = contrib.sum()
            gw = np.where(passed, dl, 0.0) / cnt                         # [M]
            Wm = mask * gw[:, None]                                       # ybar [M,q]
            mubar[s] = Wm.sum(0)
            Lbar = np.tril((Wm * sg).T @ z)                               # sum_m ebar_m z_m^T
            Pm = np.tril(L.T @ Lbar)
            Pm[np.diag_indices(q)] *= 0.5
            Xm = spla.solve_triangular(L.T, Pm, lower=False)              # L^-T P
            Sm = spla.solve_triangular(L.T, Xm.T, lower=False).T          # X L^-1
            Sigbar[s] = 0.5 * (Sm + Sm.T)
        return loss, mubar, Sigbar, chols

    def backward(self, X, p, Xs, r2, Vt, r2s, mubar, Sigbar):
        S, q, d = X.shape
        W = Sigbar + Sigbar.transpose(0, 2, 1)
        Vbar = mubar[:, :, None] * self.gamma[None, None, :] - W @ Vt     # [S,q,n]
        Kbar = Vbar @ self.Linv                                           # [S,q,n]
        g = Kbar * kern_dr2(self.kid, r2)                                 # [S,q,n]
        diff = Xs[:, :, None, :] - self.X0s[None, None, :, :]
        acc = (g[..., None] * diff).sum(2)                                # [S,q,d]
        gs = W * kern_dr2(self.kid, r2s)
        idx = np.arange(q)
        gs[:, idx, idx] = 0.0
        diffs = Xs[:, :, None, :] - Xs[:, None, :, :]
        acc += (gs[..., None] * diffs).sum(2)
        grad = 2.0 * self.amp2 * self.il * acc
        return grad[:, p:, :]

    def value_and_grad(self, acq, X, z, p=0, **kw):
        Xs, r2, Vt, mu, Sig, r2s = self.forward(X)
        loss, mubar, Sigbar, _ = self.mc(acq, mu, Sig, z, **kw)
        return loss, self.backward(X, p, Xs, r2, Vt, r2s, mubar, Sigbar)

    def closed_form(self, acq, X, level=None, beta=2.0, lower=True):
        from scipy.special import erfc
        Xs, r2, Vt, mu, Sig, r2s = self.forward(X)
        m, v = mu[:, 0], Sig[:, 0, 0]
        level = self.ymin if level is None else level
        if acq == "sr":
            l, dm, dv = m, np.ones_like(m), np.zeros_like(m)
        elif acq == "cb":
            sb = np.sqrt(beta * v)
            sgn = -1.0 if lower else 1.0
            l, dm, dv = m + sgn * sb, np.ones_like(m), sgn * beta / (2.0 * sb)
        else:
            r = level - m
            sd = np.sqrt(v)
            zz = r / sd
            pdf = np.exp(-0.5 * zz ** 2) / math.sqrt(2 * math.pi)
            cdf = 0.5 * erfc(-zz / math.sqrt(2))
            if acq == "ei":
                l, dm, dv = -(sd * pdf + r * cdf), cdf, -pdf / (2.0 * sd)
            else:
                l, dm, dv = -cdf, pdf / sd, pdf * zz / (2.0 * v)
        grad = self.backward(X, 0, Xs, r2, Vt, r2s, dm[:, None], dv[:, None, None])
        return l, grad
