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
