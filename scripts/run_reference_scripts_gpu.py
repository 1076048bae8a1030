"""The reference's UNMODIFIED scripts (scripts/run_bayesopt.py joint and pending-point greedy, scripts/
run_bayesopt_incremental.py) driving the mirror on the CUDA engine, through compat/launcher.py.  The reference checkout is
not part of this repository: MAF_REFERENCE names it (on the GPU box a transient, untracked copy under baseline/_ref/, removed
after the run).  One JSON line per episode."""
import json
import logging
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MAF_REFERENCE", os.path.join(ROOT, "baseline", "_ref", "reference"))


def make_config(**over):
    cfg = dict(seed=7, float="float64", save=0, result_dir="results", overwrite=0, task="hartmann3", loss="negative_ei",
               loss_kwargs={"num_fantasies": 128}, input_dim=3, num_initial=32, job_limit=40, parallelism=2,
               asynchronous=0, avg_runtime=60, sleep_tic=1, risk=0.0, kernel="matern52", lenscale=np.sqrt(3) / 4,
               amplitude=1.0, noise=1e-3, anisotropic=1, use_true_prior=0, num_options=None, num_to_optimize=64,
               time_limit=2.0, schedule=None, use_cpu_time=0, subroutine="sgd",
               subroutine_kwargs={"sgd": {"step_limit": 50}, "scipy": {"maxiter": 30}},
               subroutine_defaults=os.path.join(REF, "experiments/defaults/subroutines.yaml"), use_greedy=0,
               num_starts_fmin=0, num_options_fmin=1024, visualize=False, tf_logging="ERROR", end2bp=False)
    cfg.update(over)
    return types.SimpleNamespace(**cfg)


def main():
    if not os.path.isdir(os.path.join(REF, "scripts")):
        raise SystemExit("reference checkout not found at %s (set MAF_REFERENCE)" % REF)
    from maximizingacquisitionfunctions_b200 import runtime
    from maximizingacquisitionfunctions_b200.compat import launcher
    logging.basicConfig(level=logging.WARNING)
    episodes = [
        ("run_bayesopt.py joint qEI q=2 (BASELINE config 1: Hartmann-3, n_train 32 -> 40, 64 restarts, 128 samples)", "scripts/run_bayesopt.py", dict(use_greedy=0)),
        ("run_bayesopt.py pending-point greedy qEI q=2", "scripts/run_bayesopt.py", dict(use_greedy=1)),
        ("run_bayesopt_incremental.py fantasy-conditioned greedy LCB q=3, F=16", "scripts/run_bayesopt_incremental.py",
         dict(use_greedy=1, num_fantasies=16, subroutine="scipy", loss="confidence_bound", loss_kwargs={"beta": 2.0}, parallelism=3, job_limit=41)),
    ]
    for k, (name, script, over) in enumerate(episodes):
        mod = launcher.load_script(REF, script, name="maf_reference_script_%d" % k)
        mod.logger = logging.getLogger("run_experiment")
        cfg = make_config(**over)
        t0 = time.perf_counter()
        exp = mod.run_experiment(cfg)
        wall = time.perf_counter() - t0
        X = np.asarray([exp["inputs"][j] for j in sorted(exp["inputs"])])
        Y = np.asarray([exp["outputs"][j] for j in sorted(exp["outputs"])])
        eng = runtime.get_engine()
        rec = {"episode": name, "script": script, "engine": type(eng).__module__ + "." + type(eng).__name__, "library": eng.version,
               "kernel_launches": int(eng.launch_count), "observations": int(X.shape[0]), "inputs_in_unit_cube": bool(X.min() >= 0 and X.max() <= 1),
               "best_observed": float(np.min(Y)), "f_min_of_task": float(exp["f_min"]), "answers_recorded": len(exp["answer"]),
               "hyperparameter_fits": len(exp["hyperparameters"]), "wall_s": wall}
        print(json.dumps(rec), flush=True)
        assert rec["engine"].endswith("engine.Engine") and rec["kernel_launches"] > 0
        runtime.release_engines()


if __name__ == "__main__":
    main()
