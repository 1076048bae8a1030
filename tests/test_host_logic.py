"""CPU tests of the host-side logic: synthetic inputs, task objectives vs reference golden
values, restart sharding and the global arg-min record selection (incl. a world_size-2 gloo run)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from maximizingacquisitionfunctions_b200 import dist as mdist
from maximizingacquisitionfunctions_b200 import synthetic, tasks
from oracle import maf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tasks_match_reference_numpy(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_tasks.npz"))
    for name in ("hartmann3", "hartmann6", "braninhoo", "levy"):
        got = getattr(tasks, name)(g[name + "_X"])
        assert got.shape == g[name + "_Y"].shape
        assert np.abs(got - g[name + "_Y"]).max() <= 1e-14 * max(1.0, np.abs(g[name + "_Y"]).max())
        assert tasks.MINIMA[name] == float(g[name + "_min"])


def test_synthetic_inputs_agree_with_oracle_generator():
    gp, X0, y0, X, z = orc.synthetic_problem(40, 3, 2, 5, 16, seed=4)
    a = synthetic.synthetic_inputs(40, 3, 2, 5, 16, seed=4)
    for u, v in zip((X0, y0, X, z), a):
        assert np.array_equal(u, v)
    hp = synthetic.hyperparameters(3)
    assert math.isclose(float(gp.lenscales()), hp["lenscales"][0])


def test_shard_range_partitions():
    for S in (1, 7, 64, 65536):
        for W in (1, 2, 3, 8):
            spans = [mdist.shard_range(S, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pick_record_tie_break_and_nan():
    x = np.arange(6.0).reshape(2, 3)
    recs = np.array([
        np.concatenate([[0.5, 10], x.ravel()]),
        np.concatenate([[0.25, 40], (x + 1).ravel()]),
        np.concatenate([[0.25, 20], (x + 2).ravel()]),
        np.concatenate([[float("nan"), 5], x.ravel()]),
        np.concatenate([[float("inf"), -1], x.ravel()]),
    ])
    loss, idx, xb = mdist.pick_record(recs, (2, 3))
    assert (loss, idx) == (0.25, 20) and np.array_equal(xb, x + 2)
    assert mdist.local_best(np.array([np.nan, 3.0, 1.0, 1.0]), 100) == (1.0, 102)
    assert mdist.local_best(np.array([]), 5) == (float("inf"), -1)


_GLOO_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
from maximizingacquisitionfunctions_b200 import dist as mdist
d = mdist.init_process_group("gloo")
rank, ws, _ = mdist.world()
S, q, dd = 11, 2, 3
rng = np.random.RandomState(0)
losses = rng.randn(S); losses[7] = losses[2] = losses.min() - 1.0      # tie across ranks
X = rng.rand(S, q, dd)
lo, hi = mdist.shard_range(S, rank, ws)
val, gidx = mdist.local_best(losses[lo:hi], lo)
loss, idx, xb = mdist.select_global_best(val, gidx, X[gidx])
assert idx == int(np.argmin(losses)) == 2, idx
assert loss == losses[2] and np.array_equal(xb, X[2])
assert mdist.max_over_ranks(float(rank)) == ws - 1
mdist.barrier()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()
"""


def test_global_argmin_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


_SHARD_WORKER = r"""
import json, math, os, sys
import numpy as np
sys.path.insert(0, {root!r})
from maximizingacquisitionfunctions_b200 import dist as mdist, runtime
from maximizingacquisitionfunctions_b200.src import agents, losses, models
from tests.fake_engine import FakeEngine
runtime.set_engine_factory(lambda device, dtype: FakeEngine(device, dtype))
rank, ws, _ = mdist.world()
d, n = 2, 24
rng = np.random.RandomState(0)
X0 = rng.rand(n, d)
Y0 = np.sin(3 * X0.sum(-1, keepdims=True)) + 0.03 * rng.randn(n, 1)
model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(math.sqrt(d) / 4) * np.ones(d),
                                log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0)
loss_fn = losses.negative_ei(seed=2, num_fantasies=64)
agent = agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=3)
out = []
pend = None
for it in range(2):                                   # second call: one pending point, streams must still agree
    x, l = agent.suggest_inputs(2, X0, Y0, None, inputs_pend=pend, num_to_optimize=7,
                                subroutine=agent._optimize_inputs_sgd, config={{"step_limit": 6, "batch_size": 16}})
    out.append([x.tolist(), float(l), int(agent.last_argmin)])
    pend = x[:1]
x1, l1 = agent.suggest_inputs(1, X0, Y0, None, num_to_optimize=5, subroutine=agent._optimize_inputs_sgd,
                              config={{"step_limit": 4}})   # q == 1: closed form
out.append([x1.tolist(), float(l1), int(agent.last_argmin)])
if rank == 0:
    open({out!r} + ".%d" % ws, "w").write(json.dumps(out))
mdist.barrier()
"""


def test_suggest_inputs_shards_restarts_over_ranks(tmp_path):
    """agent.suggest_inputs under torchrun (world_size 2, gloo, oracle-backed engine): restarts are partitioned over
    the ranks inside the product API and the result -- suggestion, loss AND selected restart index -- is the one a
    single process finds (same host-RNG streams, same base samples, lowest-index tie-break)."""
    import json
    script = tmp_path / "shard_worker.py"
    out = str(tmp_path / "result.json")
    script.write_text(_SHARD_WORKER.format(root=ROOT, out=out))
    env = {**os.environ, "OMP_NUM_THREADS": "1"}
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    port = 29500 + ((os.getpid() + 7) % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    one, two = json.load(open(out + ".1")), json.load(open(out + ".2"))
    assert len(one) == len(two) == 3
    for (x1, l1, i1), (x2, l2, i2) in zip(one, two):
        assert i1 == i2
        assert np.array_equal(np.asarray(x1), np.asarray(x2)) and l1 == l2


def test_fast_weighted_draw_equals_numpy_legacy_choice():
    """sample_starts draws its index sets with a cached cumulative sum; the result must be the one
    RandomState.choice(n, k, p=w, replace=False) gives from the same stream (collisions and redraws included)."""
    from maximizingacquisitionfunctions_b200.src.agents.bayesian_optimization import _choice_without_replacement
    rs = np.random.RandomState(0)
    for n, k in ((5, 1), (5, 3), (6, 6), (1000, 1), (1000, 4), (17, 9)):
        w = rs.rand(n) ** 6 + 1e-16                      # heavy-tailed: collisions are frequent for small n
        w /= w.sum()
        raw = np.cumsum(w)
        cdf = raw / raw[-1]
        for seed in range(40):
            a = np.random.RandomState(seed).choice(n, k, p=w, replace=False)
            b = _choice_without_replacement(np.random.RandomState(seed), w, cdf, k)              # NumPy's redraw rounds
            c = _choice_without_replacement(np.random.RandomState(seed), w, cdf, k, raw)         # step-function redraws
            assert a.dtype.kind == b.dtype.kind and np.array_equal(a, b), (n, k, seed)
            assert np.array_equal(a, c), (n, k, seed)
