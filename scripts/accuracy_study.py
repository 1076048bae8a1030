"""Accuracy of the CUDA path against the CPU oracle at the headline size (n=4096, d=6, q=8, M=128) as a function of
the digit planes kept in the forward / reverse INT8 contraction and of where the posterior mean is computed.
Evidence script (one JSON line per case); the oracle is the checker, nothing here is a bench value.

Workloads: 'bench'  = exactly what bench.py times (uniform random candidates, qEI with level = median(y0));
           'near'   = half of the restarts 1e-3 away from training points (worst case for Sigma = K11 - V^T V);
           'smooth' = same inputs, outputs from the Hartmann-6 task + sqrt(1e-3) noise instead of white noise.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maf_oracle as orc  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-300))


def per_restart(a, b):
    a, b = np.asarray(a), np.asarray(b).reshape(np.shape(a))
    S = a.shape[0]
    return float(max(np.abs(a[s] - b[s]).max() / max(np.abs(b[s]).max(), 1e-300) for s in range(S)))


def main():
    from maximizingacquisitionfunctions_b200.engine import Engine, acq_params
    from maximizingacquisitionfunctions_b200 import tasks
    n, d, q, S, M = 4096, 6, 8, int(os.environ.get("STUDY_S", 64)), 128
    kernels = os.environ.get("STUDY_KERNELS", "matern52,squared_exponential").split(",")
    out = open(os.environ.get("STUDY_OUT", "gpurun_out/accuracy_study.jsonl"), "a")
    for kernel_id in kernels:
        for wl in os.environ.get("STUDY_WORKLOADS", "bench,near,smooth").split(","):
            gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=0, kernel_id=kernel_id)
            if wl == "near":
                X[: S // 2] = np.clip(X0[: (S // 2) * q].reshape(S // 2, q, d) + 1e-3 * np.random.RandomState(7).randn(S // 2, q, d), 0, 1)
            if wl == "smooth":
                y0 = (tasks.hartmann6(X0).reshape(n, 1) + math.sqrt(1e-3) * np.random.RandomState(1).randn(n, 1)).reshape(y0.shape)
            level = float(np.median(y0))
            ts = orc.prepare_train(gp, X0, y0)
            acqs = {"ei_median": (orc.Acq("ei", level=level), acq_params("ei", level=level)),
                    "cb": (orc.Acq("cb"), acq_params("cb")), "sr": (orc.Acq("sr"), acq_params("sr"))}
            ref = {k: orc.value_and_grad(ts, a[0], X, z) for k, a in acqs.items()}
            with torch.no_grad():
                mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
            mo, co = mo.numpy(), co.numpy()
            sig_min = float(np.diagonal(co, axis1=1, axis2=2).min())
            sig_med = float(np.median(np.diagonal(co, axis1=1, axis2=2)))
            for mu_alpha in (1, 0):
                for tight in (0, 1):
                    os.environ["MAF_MU_ALPHA"] = str(mu_alpha)
                    os.environ["MAF_VBAR_TIGHT"] = str(tight)
                    e = Engine(0)
                    try:
                        e.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
                        e.set_train(X0, y0)
                        for nf, nb in ((7, 7), (7, 6), (6, 7), (6, 6), (7, 5)):
                            if tight and nb == 7 and nf == 7 and not mu_alpha:
                                continue
                            e.set_digits(nf, nb)
                            for name, (_, prm) in acqs.items():
                                lg, gg = e.value_and_grad(prm, X, z)
                                mu, cov, _ = e.last_extras(S, q)
                                lo, go = ref[name]
                                rec = {"kernel": kernel_id, "workload": wl, "acq": name, "mu_alpha": mu_alpha, "vbar_tight": tight,
                                       "ns_fwd": nf, "ns_bwd": nb, "products": (nf * (nf + 1)) // 2 + (nb * (nb + 1)) // 2,
                                       "loss": rel(lg, lo), "loss_per_restart": per_restart(lg.reshape(S, 1), lo.reshape(S, 1)),
                                       "grad": rel(gg, go), "grad_per_restart": per_restart(gg, go),
                                       "mu": rel(mu, mo), "cov": rel(cov, co), "cov_per_restart": per_restart(cov, co),
                                       "sigma_diag_min": sig_min, "sigma_diag_median": sig_med,
                                       "loss_abs_min": float(np.abs(lo).min()), "same_argmin": int(np.argmin(lg)) == int(np.argmin(lo))}
                                out.write(json.dumps(rec) + "\n")
                                out.flush()
                    finally:
                        e.close()
    print("accuracy study done")


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("%.0f s" % (time.time() - t0))
