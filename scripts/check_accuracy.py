"""Accuracy of the CUDA path against the CPU oracle at the headline training-set size (n=4096, d=6, q=8, M=128),
for the digit configurations of the reverse contraction (MAF_OZ_NS_BWD = 6 default / 7) and for three kernels.
Evidence script (writes one JSON line per case); the oracle is the checker, nothing here is timed."""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maf_oracle as orc  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-300))


ACQS = ("cb", "sr")            # EI is identically 0 on this problem (level = min y0 of 4096 N(0,1) draws)


def main():
    from maximizingacquisitionfunctions_b200.engine import Engine, acq_params
    n, d, q, S, M = 4096, 6, 8, 32, 128
    for kernel_id in ("matern52", "squared_exponential", "matern32"):
        gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=0, kernel_id=kernel_id)
        # half of the restarts sit close to training points (small posterior variance: worst case for Sigma)
        X[: S // 2] = np.clip(X0[: (S // 2) * q].reshape(S // 2, q, d) + 1e-3 * np.random.RandomState(7).randn(S // 2, q, d), 0, 1)
        ts = orc.prepare_train(gp, X0, y0)
        ref = {}
        for acq in ACQS:
            ref[acq] = orc.value_and_grad(ts, orc.Acq(acq), X, z)
        with torch.no_grad():
            mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
        for ns_bwd in (6, 7, 0):                             # 0: FP64 DMMA contraction (no digit planes)
            os.environ["MAF_OZ_NS_BWD"] = str(ns_bwd)
            e = Engine(0)
            try:
                if ns_bwd == 0:
                    e.set_gemm_mode(0)
                e.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
                e.set_train(X0, y0)
                for acq in ACQS:
                    lg, gg = e.value_and_grad(acq_params(acq), X, z)
                    mu, cov, chol = e.last_extras(S, q)
                    lo, go = ref[acq]
                    print(json.dumps({"kernel": kernel_id, "acq": acq, "n": n, "ns_bwd": ns_bwd,
                                      "rel_err_value": relerr(lg, lo), "rel_err_grad": relerr(gg, go),
                                      "rel_err_grad_per_restart_max": float(max(
                                          np.abs(gg[s] - go[s]).max() / max(np.abs(go[s]).max(), 1e-300) for s in range(S))),
                                      "rel_err_mu": relerr(mu, mo.numpy()), "rel_err_cov": relerr(cov, co.numpy()),
                                      "same_argmin": int(np.argmin(lg)) == int(np.argmin(lo))}), flush=True)
            finally:
                e.close()
                del os.environ["MAF_OZ_NS_BWD"]


if __name__ == "__main__":
    main()
