"""Level switch of the contractions (include/maf.h, maf_set_level_switch) at the headline size: share of the restarts that
are recomputed with all digit planes, and the error against the CPU oracle per tensor and per restart, with the switch on
(default) and off.  Evidence script; the oracle is the checker."""
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
    from maximizingacquisitionfunctions_b200 import tasks
    from maximizingacquisitionfunctions_b200.engine import Engine, acq_params
    n, d, q, S, M = 4096, 6, 8, int(os.environ.get("STUDY_S", 256)), 128
    out = open(os.environ.get("STUDY_OUT", "gpurun_out/level_switch_study.jsonl"), "a")
    for kernel_id in os.environ.get("STUDY_KERNELS", "matern52,squared_exponential").split(","):
        for wl in ("bench", "near", "smooth"):
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
            for on, tol in ((1, None), (1, 1e-9), (0, None)):
                e = Engine(0)
                try:
                    e.set_level_switch(bool(on), tol)
                    e.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
                    e.set_train(X0, y0)
                    for name, (_, prm) in acqs.items():
                        e.level_switch_stats(reset=True)
                        lg, gg = e.value_and_grad(prm, X, z)
                        st = e.level_switch_stats()
                        mu, cov, _ = e.last_extras(S, q)
                        lo, go = ref[name]
                        rec = {"kernel": kernel_id, "workload": wl, "acq": name, "switch": on, "stats": st,
                               "share_fwd": st["recomputed_forward"] / max(st["tested_forward"], 1),
                               "share_bwd": st["recomputed_reverse"] / max(st["tested_reverse"], 1),
                               "loss": rel(lg, lo), "loss_per_restart": per_restart(lg.reshape(S, 1), lo.reshape(S, 1)),
                               "grad": rel(gg, go), "grad_per_restart": per_restart(gg, go),
                               "mu": rel(mu, mo), "mu_per_restart": per_restart(mu, mo), "cov": rel(cov, co),
                               "cov_per_restart": per_restart(cov, co), "same_argmin": int(np.argmin(lg)) == int(np.argmin(lo))}
                        out.write(json.dumps(rec) + "\n")
                        out.flush()
                        print(kernel_id[:3], wl, name, "switch", on, "tol", st["tol"], "share f/b %.3f %.3f" % (rec["share_fwd"], rec["share_bwd"]),
                              "loss/r %.1e grad/r %.1e cov/r %.1e mu/r %.1e" % (rec["loss_per_restart"], rec["grad_per_restart"],
                                                                                 rec["cov_per_restart"], rec["mu_per_restart"]), flush=True)
                finally:
                    e.close()


if __name__ == "__main__":
    main()
