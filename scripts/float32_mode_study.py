"""dtype='float32' mode (the reference's float32 runs: jitter 1e-4, parity bar 1e-4): accuracy against the float64 oracle at
the headline size for 5 / 4 / 3 digit planes per contraction (15 / 10 / 6 exact products).  One JSON line per case."""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maf_oracle as orc  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-300))


def per_restart(a, b):
    a, b = np.asarray(a), np.asarray(b).reshape(np.shape(a))
    return float(max(np.abs(a[s] - b[s]).max() / max(np.abs(b[s]).max(), 1e-300) for s in range(a.shape[0])))


def main():
    from maximizingacquisitionfunctions_b200.engine import Engine, acq_params
    n, d, q, S, M = 4096, 6, 8, 64, 128
    for kernel_id in ("matern52", "squared_exponential"):
        for wl in ("bench", "near"):
            gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=0, kernel_id=kernel_id)
            gp.jitter = 1e-4
            if wl == "near":
                X[: S // 2] = np.clip(X0[: (S // 2) * q].reshape(S // 2, q, d) + 1e-3 * np.random.RandomState(7).randn(S // 2, q, d), 0, 1)
            level = float(np.median(y0))
            ts = orc.prepare_train(gp, X0, y0)
            lo, go = orc.value_and_grad(ts, orc.Acq("ei", level=level), X, z)
            with torch.no_grad():
                mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
            e = Engine(0, "float32")
            try:
                e.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-4)
                e.set_train(X0, y0)
                for ns in (5, 4, 3):
                    e.set_digits(ns, ns)
                    lg, gg = e.value_and_grad(acq_params("ei", level=level), X, z)
                    mu, cov, _ = e.last_extras(S, q)
                    print(json.dumps({"kernel": kernel_id, "workload": wl, "planes": ns, "products": ns * (ns + 1) // 2,
                                      "loss": rel(lg, lo), "loss_per_restart": per_restart(lg.reshape(S, 1), lo.reshape(S, 1)),
                                      "grad": rel(gg, go), "grad_per_restart": per_restart(gg, go), "mu": rel(mu, mo.numpy()),
                                      "cov": rel(cov, co.numpy()), "cov_per_restart": per_restart(cov, co.numpy()),
                                      "same_argmin": int(np.argmin(lg)) == int(np.argmin(lo))}), flush=True)
            finally:
                e.close()


if __name__ == "__main__":
    main()
