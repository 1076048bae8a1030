"""Own blocked Cholesky / triangular inverse (csrc/factor.cuh) against NumPy/SciPy, with timings of maf_set_train."""
import math, os, sys, time
import numpy as np
import scipy.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200.engine import Engine

def k00(X0, ls, noise, jitter):
    Z = X0 / ls
    r2 = np.maximum(((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1), 0.0) if len(X0) <= 2048 else None
    if r2 is None:
        s = (Z * Z).sum(1)
        r2 = np.maximum(s[:, None] + s[None, :] - 2 * Z @ Z.T, 0.0)
    u = 5.0 * np.maximum(r2, np.finfo(float).eps); r = np.sqrt(u)
    K = (1 + r + u / 3) * np.exp(-r)
    K[np.diag_indices_from(K)] += noise + jitter
    return K

for n in (100, 384, 1000, 4096, 8300):
    rng = np.random.RandomState(n); X0 = rng.rand(n, 4); y0 = rng.randn(n)
    e = Engine(0); e.set_model("matern52", np.full(4, 0.5), 1.0, 1e-3, 0.0, 1e-6)
    t0 = time.perf_counter(); e.set_train(X0, y0); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); e.set_train(X0, y0 + 1e-3); t2 = time.perf_counter() - t0
    L0, Linv, gam = e.factors()
    msg = "n=%d set_train %.1f ms (first) / %.1f ms" % (n, t1 * 1e3, t2 * 1e3)
    if n <= 4096:
        K = k00(X0, 0.5, 1e-3, 1e-6)
        Lr = spla.cholesky(K, lower=True)
        msg += "  |L-Lref|max=%.2e  |LL^T-K|/|K|=%.2e" % (np.abs(L0 - Lr).max(), np.abs(L0 @ L0.T - K).max() / np.abs(K).max())
    msg += "  |Linv L - I|max=%.2e  upper(L)=%g upper(Linv)=%g" % (np.abs(Linv @ L0 - np.eye(n)).max(), np.abs(np.triu(L0, 1)).max(), np.abs(np.triu(Linv, 1)).max())
    print(msg, flush=True)
    e.close()
