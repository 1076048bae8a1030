"""CPU: the reference's UNMODIFIED scripts/run_bayesopt.py drives the mirror end to end (host logic on
the test-only oracle engine).  Needs the reference checkout, so it only runs in the build container."""
import logging
import os
import types

import numpy as np
import pytest

from maximizingacquisitionfunctions_b200 import runtime
from tests.fake_engine import FakeEngine

REF = os.environ.get("MAF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="reference checkout not present")


@pytest.fixture(autouse=True)
def fake_engine():
    runtime.set_engine_factory(lambda device, dtype: FakeEngine(device, dtype))
    yield
    runtime.set_engine_factory(None)


def make_config(**over):
    cfg = dict(seed=7, float="float64", save=0, result_dir="results", overwrite=0, task="hartmann3", loss="negative_ei",
               loss_kwargs={"num_fantasies": 64}, input_dim=3, num_initial=4, job_limit=8, parallelism=2,
               asynchronous=0, avg_runtime=60, sleep_tic=1, risk=0.0, kernel="matern52", lenscale=np.sqrt(3) / 4,
               amplitude=1.0, noise=1e-3, anisotropic=1, use_true_prior=0, num_options=None, num_to_optimize=None,
               time_limit=0.05, schedule=None, use_cpu_time=0, subroutine="sgd",
               subroutine_kwargs={"sgd": {"step_limit": 6}, "scipy": {"maxiter": 4}},
               subroutine_defaults=os.path.join(REF, "experiments/defaults/subroutines.yaml"), use_greedy=0,
               num_starts_fmin=0, num_options_fmin=256, visualize=False, tf_logging="ERROR", end2bp=False)
    cfg.update(over)
    return types.SimpleNamespace(**cfg)


@pytest.mark.parametrize("use_greedy", [0, 1])
def test_run_bayesopt_script_unmodified(use_greedy):
    from maximizingacquisitionfunctions_b200.compat import launcher
    mod = launcher.load_script(REF, "scripts/run_bayesopt.py")
    mod.logger = logging.getLogger("run_experiment")
    cfg = make_config(use_greedy=use_greedy)
    exp = mod.run_experiment(cfg)
    X = np.asarray([exp["inputs"][k] for k in sorted(exp["inputs"])])       # job id -> input
    Y = np.asarray([exp["outputs"][k] for k in sorted(exp["outputs"])])
    assert X.shape == (8, 3) and Y.shape[0] == 8
    assert X.min() >= 0 and X.max() <= 1
    assert len(exp["answer"]) >= 2 and len(exp["hyperparameters"]) >= 1
    assert np.isfinite(exp["f_min"])


def test_run_bayesopt_incremental_script_unmodified():
    """scripts/run_bayesopt_incremental.py: fantasy-conditioned greedy batches (rank-4 observation sets),
    incl. its direct use of get_or_create_ref / get_or_create_node / sess.run (:150-163)."""
    from maximizingacquisitionfunctions_b200.compat import launcher
    mod = launcher.load_script(REF, "scripts/run_bayesopt_incremental.py", name="maf_reference_incremental")
    mod.logger = logging.getLogger("run_experiment")
    cfg = make_config(use_greedy=1, num_fantasies=8, subroutine="scipy", loss="confidence_bound",
                      loss_kwargs={"beta": 2.0}, parallelism=3, job_limit=10)
    exp = mod.run_experiment(cfg)
    X = np.asarray([exp["inputs"][k] for k in sorted(exp["inputs"])])
    assert X.shape == (10, 3) and X.min() >= 0 and X.max() <= 1
