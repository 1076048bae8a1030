"""BASELINE config 4 as specified, through the product API (the agent mirror) on one GPU: Levy d=6, n=2048, 8192 restarts,
batch of 16 built greedily in the reference's two ways --
  (i)  pending-point greedy qLCB (scripts/run_bayesopt.py:137-176): 16 calls suggest_inputs(num_suggest=1, inputs_pend=picks),
       pool size grows 1 -> 16, Adam on the Monte-Carlo loss for q > 1 (128 base samples per step), L-BFGS closed form at q = 1;
  (ii) incremental / fantasy-conditioned greedy (scripts/run_bayesopt_incremental.py:138-218): 16 calls suggest_inputs(1) on
       rank-4 observations (F = 64 fantasies, n grows 2048 -> 2063), closed-form LCB, device L-BFGS.
One JSON line per greedy step (wall seconds of the call, device ms per value+gradient pass where timed) and a summary.
MAF_PEND_DEDUPE=0 repeats (i) with every restart contracting its own copy of the pending rows (the round-1 behaviour)."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200 import runtime, tasks  # noqa: E402
from maximizingacquisitionfunctions_b200.src import agents, losses, models  # noqa: E402


def build(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X0 = rng.rand(n, d)
    Y0 = tasks.levy(X0) + math.sqrt(1e-3) * rng.randn(n, 1)
    Y0 = (Y0 - Y0.mean()) / Y0.std()
    model = models.gaussian_process(kernel_id="matern52", seed=seed + 1, log_lenscales=np.log(math.sqrt(d) / 4) * np.ones(d),
                                    log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0)
    loss_fn = losses.confidence_bound(seed=seed + 2, beta=2.0, lower=True, num_fantasies=1024)
    return agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=seed + 3), X0, Y0


def main():
    n, d, S, Q = 2048, 6, int(os.environ.get("C4_S", 8192)), int(os.environ.get("C4_Q", 16))
    steps = int(os.environ.get("C4_ADAM_STEPS", 8))
    out = []
    mode = os.environ.get("C4_MODE", "both")
    if mode in ("both", "pending"):
        agent, X0, Y0 = build(n, d)
        eng = agent.model.bind(X0, Y0)
        pend, total = None, 0.0
        for k in range(Q):
            t0 = time.perf_counter()
            x, l = agent.suggest_inputs(1, X0, Y0, None, inputs_pend=pend, num_to_optimize=S,
                                        config={"step_limit": steps, "batch_size": 128, "maxiter": 20})
            dt = time.perf_counter() - t0
            total += dt
            pend = x if pend is None else np.vstack([pend, x])
            # device time of ONE value+gradient pass at this pool size (pending rows included), same restarts
            pools = np.concatenate([np.broadcast_to(pend[:-1], (S, k, d)), np.random.RandomState(k).rand(S, 1, d)], axis=1) if k else np.random.RandomState(k).rand(S, 1, d)
            z = None if k == 0 else np.random.RandomState(100 + k).randn(128, k + 1)
            prm = agent.loss_fn.acq_params(outputs_old=Y0)
            eng.value_and_grad(prm, pools, z, p_pending=k)
            t1 = time.perf_counter()
            for _ in range(3):
                eng.value_and_grad(prm, pools, z, p_pending=k)
            ms = (time.perf_counter() - t1) / 3 * 1e3
            rec = {"mode": "pending-point greedy qLCB", "step": k + 1, "pool_size": k + 1, "pending": k, "suggest_inputs_wall_s": dt,
                   "value_grad_pass_ms_host_api": ms, "loss": float(l), "dedupe": os.environ.get("MAF_PEND_DEDUPE", "1")}
            print(json.dumps(rec), flush=True)
            out.append(rec)
        print(json.dumps({"mode": "pending-point greedy qLCB", "summary": True, "batch": Q, "restarts": S, "n": n,
                          "total_wall_s": total, "sum_value_grad_pass_ms": sum(r["value_grad_pass_ms_host_api"] for r in out),
                          "dedupe": os.environ.get("MAF_PEND_DEDUPE", "1"), "picks_min_pairwise_distance": float(min(
                              np.linalg.norm(pend[i] - pend[j]) for i in range(len(pend)) for j in range(i)) if len(pend) > 1 else 0.0)}), flush=True)
        runtime.release_engines()
    if mode in ("both", "incremental"):
        F = int(os.environ.get("C4_F", 64))
        agent, X0, Y0 = build(n, d, seed=1)
        Xf, Yf = X0, Y0
        total, picks = 0.0, []
        for k in range(Q):
            t0 = time.perf_counter()
            x, l = agent.suggest_inputs(1, Xf, Yf, None, num_to_optimize=S, config={"maxiter": 20})
            t_suggest = time.perf_counter() - t0
            # fantasise F outcomes at the pick and pad the observation set (run_bayesopt_incremental.py:150-192)
            if Xf.ndim == 2:
                fant = agent.model.draw_samples(F, x, Xf, Yf)                                   # [F, 1]
                Xf = np.tile(np.vstack([Xf, x])[None, None], [F, 1, 1, 1])
                Yf = np.dstack([np.tile(Yf[None, None], [F, 1, 1, 1]), fant[:, None, None]])
            else:
                fant = agent.model.draw_samples(1, x, Xf, Yf)                                   # [F, 1, 1, 1]
                Xf = np.dstack([Xf, np.tile(x[None, None], [F, 1, 1, 1])])
                Yf = np.dstack([Yf, fant])
            dt = time.perf_counter() - t0
            total += dt
            picks.append(x[0])
            rec = {"mode": "incremental greedy LCB (F=%d)" % F, "step": k + 1, "n_obs": int(np.shape(Xf)[-2]) - 1,
                   "suggest_inputs_wall_s": t_suggest, "step_wall_s": dt, "loss": float(l)}
            print(json.dumps(rec), flush=True)
        print(json.dumps({"mode": "incremental greedy LCB (F=%d)" % F, "summary": True, "batch": Q, "restarts": S, "n": n,
                          "total_wall_s": total}), flush=True)
        runtime.release_engines()


if __name__ == "__main__":
    main()
