#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE's own code.

Runs only in the build container (needs /root/reference); the fixtures it writes
(tests/golden/*.npz) travel with the repo, the reference does not.

What can be taken from the reference without TensorFlow (absent here; TF 1.x has
no Python 3.12 build) is its NumPy/SciPy side:
  * tests/test_models.py:30-87   -- the in-file GP posterior oracle
    (``squared_exponential``, ``jitter_cholesky``, ``_gp_posterior``,
    ``gp_posterior``) on the exact configuration of ``test_gp_predict``
    (:229-256): SE kernel, l=0.1, a=1.2, noise=1e-3, num_new=64, num_old=128,
    d=4, set_size=4, rank_new x rank_old in {2,3}^2.
  * tests/test_utils.py:43-67    -- ``pdist2`` truth = scipy cdist, N=256,
    M=128, D=32, seed 1234.
  * tests/test_qUCB.py:47-53     -- ``test_1ucb`` folded-normal identity.
  * tasks/*.py ``_numpy``        -- objective values used to build the BASELINE
    configs (hartmann3/6, braninhoo, levy).
  * src/third_party/sobol_lib.py, src/core/module.py RNG semantics (restated
    stream checked against a literal RandomState(seed+k) sequence).

``tensorflow`` and ``matplotlib`` are replaced by inert MagicMock modules ONLY so
that the reference files import; no mocked attribute takes part in any number
written below.
"""
import importlib
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("MAF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    for name in ["tensorflow", "tensorflow.contrib", "tensorflow.contrib.opt",
                 "tensorflow.python", "tensorflow.python.client",
                 "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "tests"))
    import warnings
    warnings.simplefilter("ignore")
    tm = importlib.import_module("test_models")
    tq = importlib.import_module("test_qUCB")
    tasks = importlib.import_module("tasks")
    return tm, tq, tasks


def main():
    tm, tq, tasks = _import_reference()
    from scipy.spatial.distance import cdist

    # ---- 1. GP posterior (reference tests/test_models.py)
    post = {}
    hyp = dict(log_lenscales=np.log(0.1), log_amplitude=np.log(1.2), log_noise=np.log(1e-3))
    for seed in (11, 12):
        for rank_new in (2, 3):
            for rank_old in (2, 3):
                rng = np.random.RandomState(seed)
                shape_old = (rank_old - 2) * [4] + [128]
                shape_new = (rank_new - 2) * [4] + [64]
                X1 = rng.rand(*shape_new, 4)
                X0 = rng.rand(*shape_old, 4)
                Y0 = rng.rand(*shape_old, 1)
                for mean in (None, 0.25):
                    m, c = tm.gp_posterior(X1, X0, Y0, jitter=1e-6, mean=mean, **hyp)
                    key = "s{}_rn{}_ro{}_m{}".format(seed, rank_new, rank_old, "emp" if mean is None else "c")
                    post[key + "_X1"], post[key + "_X0"], post[key + "_Y0"] = X1, X0, Y0
                    post[key + "_mean"], post[key + "_cov"] = np.asarray(m), np.asarray(c)
    np.savez_compressed(os.path.join(HERE, "ref_gp_posterior.npz"), **post)

    # ---- 2. pdist2 (reference tests/test_utils.py)
    rng = np.random.RandomState(1234)
    X = rng.randn(256, 32)
    Y = rng.randn(128, 32)
    np.savez_compressed(os.path.join(HERE, "ref_pdist2.npz"), X=X, Y=Y,
                        XX=cdist(X, X, metric="sqeuclidean"),
                        XY=cdist(X, Y, metric="sqeuclidean"))

    # ---- 3. folded-normal identity (reference tests/test_qUCB.py:47-53)
    rvs = np.random.RandomState(5).randn(200000)
    rows = []
    for mu, var, beta in [(0.3, 2.0, 2.0), (-1.0, 0.04, 1.0), (0.0, 1.0, 3.0)]:
        mc, cf = tq.test_1ucb(mu, var, beta=beta, rvs=rvs)
        rows.append([mu, var, beta, mc, cf])
    np.savez_compressed(os.path.join(HERE, "ref_1ucb.npz"), rvs_seed=5, n=200000, rows=np.asarray(rows))

    # ---- 4. task values (reference tasks/*._numpy)
    tv = {}
    for name, d in [("hartmann3", 3), ("hartmann6", 6), ("braninhoo", 2), ("levy", 6)]:
        cls = getattr(tasks, name)
        task = cls(d) if name == "levy" else cls()
        Xq = np.random.RandomState(21 + d).rand(64, d)
        tv[name + "_X"] = Xq
        tv[name + "_Y"] = task._numpy(Xq)
        tv[name + "_min"] = np.asarray(task.minimum)
    np.savez_compressed(os.path.join(HERE, "ref_tasks.npz"), **tv)

    # ---- 5. host RNG stream (reference src/core/module.py:29-42,131-144)
    core = importlib.import_module("src.core")
    mod = core.module(seed=77)
    draws = [mod.rng.randn(3, 2), mod.rng.rand(4), mod.rng.randn(2, 2)]
    mod2 = core.module(seed=2 ** 31 - 3)
    wrap = [mod2.rng.rand(2) for _ in range(4)]
    np.savez_compressed(os.path.join(HERE, "ref_rng.npz"), d0=draws[0], d1=draws[1], d2=draws[2],
                        wrap=np.asarray(wrap), final_seed=np.asarray([mod.seed, mod2.seed]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
