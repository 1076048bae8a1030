"""TEST-ONLY stand-in for engine.Engine backed by the CPU oracle, so that the host-side logic of
the ``src`` mirror (agents / losses / models) can be exercised without a GPU.  Never imported by
the product package."""
import math

import numpy as np
import torch

from oracle import maf_oracle as orc

_KIND = {0: "ei", 1: "pi", 2: "sr", 3: "cb"}


class FakeEngine:
    def __init__(self, device=0, dtype="float64"):
        self.device, self.dtype = device, dtype
        self.launch_count = 0
        self._last = None

    def close(self):
        pass

    def set_model(self, kernel_id, lenscales, amplitude, noise, mean, jitter):
        self.gp = orc.GP(log_lenscales=np.log(np.asarray(lenscales, dtype=np.float64)), log_amplitude=math.log(amplitude),
                         log_noise=None if (noise is None or noise <= 0.0) else math.log(noise), mean=mean, kernel_id=kernel_id,
                         jitter=jitter)
        self.d = len(np.atleast_1d(lenscales))

    def set_train(self, X0, y0):
        self.ts = orc.prepare_train(self.gp, X0, y0)
        self.n = len(X0)
        self.F, self._fY, self._fts = 0, None, None      # like maf_set_train: the fantasies are dropped

    def factors(self):
        L = self.ts.chol.numpy()
        return L, np.linalg.inv(L), None

    def gp_nll(self):
        gp = self.gp
        d = self.d
        ll = torch.as_tensor(np.broadcast_to(np.asarray(gp.log_lenscales, dtype=np.float64), (d,)).copy()).requires_grad_(True)
        la = torch.tensor(float(gp.log_amplitude), dtype=orc.DT, requires_grad=True)
        ln = None if gp.log_noise is None else torch.tensor(float(gp.log_noise), dtype=orc.DT, requires_grad=True)
        mean_is_free = gp.mean is not None
        mm = torch.tensor(float(gp.mean) if mean_is_free else 0.0, dtype=orc.DT, requires_grad=True)
        g2 = orc.GP(log_lenscales=ll, log_amplitude=la, log_noise=ln, mean=mm if mean_is_free else None,
                    kernel_id=gp.kernel_id, jitter=gp.jitter)
        nll = -orc.log_likelihood_data(g2, self.ts.X0, self.ts.Y0)
        nll.backward()
        grad = np.zeros(d + 3)
        grad[:d] = ll.grad.numpy()
        grad[d] = la.grad.item()
        grad[d + 1] = 0.0 if ln is None else ln.grad.item()
        grad[d + 2] = mm.grad.item() if (mean_is_free and mm.grad is not None) else 0.0
        return float(nll.item()), grad

    def _acq(self, prm):
        t = None if math.isnan(prm.temperature) else prm.temperature
        lvl = None if math.isnan(prm.level) else prm.level
        return orc.Acq(_KIND[prm.acq], temperature=t, beta=prm.beta, lower=bool(prm.lower), level=lvl)

    def posterior(self, X, full_cov=True):
        with torch.no_grad():
            m, c = orc.posterior(self.ts, torch.as_tensor(X, dtype=orc.DT), full_cov=full_cov)
        return m.numpy(), c.numpy()

    def value_and_grad(self, prm, X, z=None, p_pending=0, weights=None, incumbents=None, need_grad=True,
                       out_loss=None, out_grad=None):
        X = np.asarray(X, dtype=np.float64)
        pend = X[0, :p_pending] if p_pending else None
        zz = np.zeros((1, 1)) if z is None else z
        losses, grad = orc.value_and_grad(self.ts, self._acq(prm), X[:, p_pending:], zz, X_pend=pend,
                                          weights=weights, incumbents=incumbents, shard_limit=2 ** 20)
        with torch.no_grad():
            Xt = torch.as_tensor(X)
            if X.shape[1] == 1 and z is None:
                m, v = orc.posterior(self.ts, Xt, full_cov=False)
                self._last = (m.numpy(), v.numpy(), None)
            else:
                m, c = orc.posterior(self.ts, Xt, full_cov=True)
                self._last = (m.numpy(), c.numpy(), orc.jitter_cholesky(c, self.gp.jitter).numpy())
        self._loss = losses
        return losses, (grad if need_grad else None)

    def mc_from_moments(self, prm, mu, cov, z, weights=None, incumbents=None, need_grad=False):
        cov = torch.as_tensor(np.asarray(cov, dtype=np.float64))
        S, q, _ = cov.shape
        mu = torch.as_tensor(np.asarray(mu, dtype=np.float64)).reshape(S, q, 1)
        chol = orc.jitter_cholesky(torch.tril(cov) + torch.tril(cov, -1).transpose(-1, -2), self.gp.jitter)
        y0 = getattr(self, "ts", None)
        y0 = y0.Y0 if y0 is not None else torch.zeros(1, 1, dtype=orc.DT)
        w = None if weights is None else torch.as_tensor(weights)
        inc = None if incumbents is None else torch.as_tensor(incumbents)
        loss = orc.mc_loss(self._acq(prm), mu, chol, torch.as_tensor(np.asarray(z, dtype=np.float64)), y0, w, inc)
        return loss.numpy(), None, None, chol.numpy()

    def set_fantasies(self, Y):
        self.F = len(Y)
        self._fY = np.asarray(Y, dtype=np.float64)
        self._fts = [orc.prepare_train(self.gp, self.ts.X0.numpy(), y) for y in np.asarray(Y)]

    def fantasies_value_and_grad(self, prm, X, need_grad=True):
        if not getattr(self, "F", 0):
            raise RuntimeError("maf_set_fantasies must be called first")
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.d)
        acq = self._acq(prm)
        Xt = torch.as_tensor(X[:, None, :]).clone().requires_grad_(True)
        losses = orc.fantasy_closed_form_losses(self.gp, self.ts.X0.numpy(), self._fY, acq, Xt)
        if need_grad:
            losses.sum().backward()
            return losses.detach().numpy(), Xt.grad.numpy()[:, 0, :]
        return losses.detach().numpy(), None

    def fantasies_posterior(self, X):
        if not getattr(self, "F", 0):
            raise RuntimeError("maf_set_fantasies must be called first")
        X = np.asarray(X, dtype=np.float64).reshape(-1, 1, self.d)
        mus, var = [], None
        with torch.no_grad():
            for ts in self._fts:
                m, v = orc.posterior(ts, torch.as_tensor(X), full_cov=False)
                mus.append(m.numpy().reshape(-1))
                var = v.numpy().reshape(-1)
        return np.asarray(mus), var

    def lbfgs_run(self, prm, X, z=None, p_pending=0, fantasies=False, maxiter=1000, ftol=1e6 * np.finfo("float64").eps,
                  pgtol=1e-5, time_limit=float("inf"), want_trace=True):
        """TEST-ONLY restatement of csrc/lbfgs.cuh (per-restart box-constrained L-BFGS, all restarts in lock step) on the
        oracle's value + gradient, so that the host logic above it can be compared engine against engine."""
        LB_M = 10
        X = np.array(X, dtype=np.float64)
        S, q, d = X.shape
        p = int(p_pending)
        D = (q - p) * d

        def evaluate(Xp):
            if fantasies:
                l, g = self.fantasies_value_and_grad(prm, Xp[:, 0, :])
                return np.asarray(l), np.asarray(g).reshape(S, D)
            l, g = self.value_and_grad(prm, Xp, z, p_pending=p)
            return np.asarray(l).reshape(S), np.asarray(g).reshape(S, D)

        x = np.zeros((S, D)); f = np.zeros(S); g = np.zeros((S, D)); dirn = np.zeros((S, D)); step = np.zeros(S)
        HS = np.zeros((S, LB_M, D)); HY = np.zeros((S, LB_M, D)); rho = np.zeros((S, LB_M))
        hcount = np.zeros(S, dtype=int); hpos = np.zeros(S, dtype=int); status = np.zeros(S, dtype=int)
        moved_len = np.zeros(S)
        trace, evals, active = [], 0, S
        held = lambda xv, gv: ((xv <= 0.0) & (gv > 0.0)) | ((xv >= 1.0) & (gv < 0.0))
        for it in range(int(maxiter)):
            ft, gt = evaluate(X)
            evals += 1
            trace.append(ft.copy())
            Xt = X[:, p:, :].reshape(S, D)                      # view of the trainable rows
            for s in range(S):
                first = it == 0
                if not first and status[s] != 0:
                    continue
                xt, fn, gn = Xt[s], ft[s], gt[s]
                if not first:
                    gdx = float(np.dot(g[s], xt - x[s]))
                    if not (fn == fn and fn <= f[s] + 1e-4 * gdx):
                        step[s] *= 0.5
                        v = np.clip(x[s] + step[s] * dirn[s], 0.0, 1.0)
                        moved = bool(np.any(v != x[s]))
                        Xt[s] = v
                        if step[s] < 1e-12 or not moved:
                            Xt[s] = x[s]
                            status[s] = 3
                            active -= 1
                        continue
                    sv, yv = xt - x[s], gn - g[s]
                    sy, ss, yy = float(np.dot(sv, yv)), float(np.dot(sv, sv)), float(np.dot(yv, yv))
                    slot = hpos[s]
                    HS[s, slot], HY[s, slot] = sv, yv
                    if sy > 1e-10 * math.sqrt(ss * yy) and sy > 0.0:
                        rho[s, slot] = 1.0 / sy
                        hpos[s] = (slot + 1) % LB_M
                        hcount[s] = min(hcount[s] + 1, LB_M)
                    moved_len[s] = math.sqrt(ss)
                    if f[s] - fn <= ftol * max(abs(f[s]), abs(fn), 1.0):
                        status[s] = 2
                f[s] = fn
                x[s] = xt.copy()
                g[s] = gn.copy()
                pg = x[s] - np.clip(x[s] - g[s], 0.0, 1.0)
                if status[s] == 0 and np.abs(pg).max() <= pgtol:
                    status[s] = 1
                if status[s] != 0:
                    active -= 1
                    continue
                hc, hp = hcount[s], hpos[s]
                hd = held(x[s], g[s])
                dv = np.where(hd, 0.0, g[s])
                al = np.zeros(LB_M)
                for u in range(hc):
                    k = (hp - 1 - u + 2 * LB_M) % LB_M
                    a = rho[s, k] * float(np.dot(HS[s, k], dv))
                    al[u] = a
                    dv = dv - a * HY[s, k]
                if hc > 0:
                    k = (hp - 1 + LB_M) % LB_M
                    dv = dv * (1.0 / (rho[s, k] * float(np.dot(HY[s, k], HY[s, k]))))
                for u in range(hc - 1, -1, -1):
                    k = (hp - 1 - u + 2 * LB_M) % LB_M
                    b = rho[s, k] * float(np.dot(HY[s, k], dv))
                    dv = dv + (al[u] - b) * HS[s, k]
                dv = np.where(hd, 0.0, -dv)
                gd = float(np.dot(g[s], dv))
                gnorm = float(np.sum(np.where(hd, 0.0, g[s] * g[s])))
                if not gd < 0.0:
                    hcount[s] = 0
                    dv = np.where(hd, 0.0, -g[s])
                a = 1.0
                if hcount[s] == 0:
                    gn2 = math.sqrt(max(gnorm, 1e-300))
                    L = min(1.0, gn2) if first else min(8.0 * moved_len[s], math.sqrt(float(D)))
                    a = L / gn2
                step[s], dirn[s] = a, dv
                v = np.clip(x[s] + a * dv, 0.0, 1.0)
                if not np.any(v != x[s]):
                    status[s] = 1
                    active -= 1
                Xt[s] = v
            X[:, p:, :] = Xt.reshape(S, q - p, d)
            if active <= 0:
                break
        X[:, p:, :] = x.reshape(S, q - p, d)
        return X, f.copy(), (np.asarray(trace) if want_trace else None), {"evaluations": evals, "not_converged": max(active, 0)}

    def last_extras(self, S, q, want_chol=True):
        return self._last

    def argmin_last(self):
        i = int(np.nanargmin(self._loss))
        return i, float(self._loss[i])

    def adam_begin(self, prm, X, p_pending=0, lr=25e-3, beta1=0.9, beta2=0.999, eps=1e-8, method="adam", momentum=0.9):
        self._ax = np.array(X, dtype=np.float64)
        self._ap, self._aprm = p_pending, prm
        self._ast = orc.adam_init(self._ax[:, p_pending:].shape, lr=lr, beta1=beta1, beta2=beta2, eps=eps)
        self._amethod, self._amom, self._alr, self._at = method, (momentum if method == "momentum" else 0.0), lr, 0
        self._aacc = np.zeros_like(self._ax[:, p_pending:])

    def adam_steps(self, z_steps, want_trace=False, steps=None):
        trace = []
        p = self._ap
        if z_steps is None:
            z_steps = [None] * steps
        for z in z_steps:
            losses, g = self.value_and_grad(self._aprm, self._ax, z, p_pending=p)
            trace.append(losses)
            if self._amethod == "adam":
                self._ax[:, p:] = orc.adam_step(self._ax[:, p:], g, self._ast)
            else:                                         # same update as oracle.sgd_run
                self._aacc = self._amom * self._aacc + g
                self._ax[:, p:] = np.clip(self._ax[:, p:] - self._alr * float(self._at + 1) ** -0.7 * self._aacc, 0.0, 1.0)
            self._at += 1
        return np.asarray(trace) if want_trace else None

    def adam_peek(self):
        return self._ax.copy()

    def adam_end(self):
        return self._ax.copy()
