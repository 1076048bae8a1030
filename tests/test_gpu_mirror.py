"""GPU: the ``src`` mirror end to end on the CUDA engine vs the same host logic on the test-only
oracle-backed engine (identical seeds => identical host-RNG streams, starts and base samples)."""
import math

import numpy as np
import pytest

from maximizingacquisitionfunctions_b200 import runtime, tasks
from maximizingacquisitionfunctions_b200.src import agents, losses, models
from tests.fake_engine import FakeEngine

pytestmark = pytest.mark.gpu


def build(d, n, task, seed, loss, **loss_kwargs):
    rng = np.random.RandomState(seed)
    X0 = rng.rand(n, d)
    Y0 = getattr(tasks, task)(X0) + math.sqrt(1e-3) * rng.randn(n, 1)
    model = models.gaussian_process(kernel_id="matern52", seed=seed + 1, log_lenscales=np.log(math.sqrt(d) / 4) * np.ones(d),
                                    log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0)
    loss_fn = getattr(losses, loss)(seed=seed + 2, **loss_kwargs)
    return agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=seed + 3), X0, Y0


def run_both(fn):
    runtime.set_engine_factory(None)
    runtime.release_engines()
    try:
        gpu = fn()
    finally:
        runtime.release_engines()
    runtime.set_engine_factory(lambda device, dtype: FakeEngine(device, dtype))
    try:
        ref = fn()
    finally:
        runtime.set_engine_factory(None)
    return gpu, ref


def test_c1_hartmann3_qei_joint_sgd():
    """BASELINE config 1: Hartmann-3, qEI q=2, n_train=32, 64 restarts, 128 MC samples per step."""
    def flow():
        agent, X0, Y0 = build(3, 32, "hartmann3", 0, "negative_ei")
        agent.loss_fn.num_fantasies = 256
        return agent.suggest_inputs(2, X0, Y0, None, num_to_optimize=64, subroutine=agent._optimize_inputs_sgd,
                                    config={"step_limit": 25, "batch_size": 128})
    (xg, lg), (xr, lr) = run_both(flow)
    assert xg.shape == (2, 3)
    assert np.abs(xg - xr).max() < 1e-6 and abs(lg - lr) < 1e-8 * max(1.0, abs(lr))


@pytest.mark.parametrize("loss,kw", [("negative_pi", {"temperature": 0.01}), ("simple_regret", {}),
                                     ("confidence_bound", {"beta": 2.0})])
def test_c2_branin_q4(loss, kw):
    def flow():
        agent, X0, Y0 = build(2, 256, "braninhoo", 1, loss, **kw)
        agent.loss_fn.num_fantasies = 256
        return agent.suggest_inputs(4, X0, Y0, None, num_to_optimize=32, subroutine=agent._optimize_inputs_sgd,
                                    config={"step_limit": 10, "batch_size": 256})
    (xg, lg), (xr, lr) = run_both(flow)
    assert np.abs(xg - xr).max() < 1e-6 and abs(lg - lr) < 1e-8 * max(1.0, abs(lr))


def test_greedy_pending_and_scipy_and_answers():
    def flow():
        agent, X0, Y0 = build(3, 48, "hartmann3", 2, "negative_ei")
        agent.loss_fn.num_fantasies = 128
        pend = None
        picks = []
        for _ in range(3):                 # scripts/run_bayesopt.py:137-176 (pending-point greedy)
            x, l = agent.suggest_inputs(1, X0, Y0, None, inputs_pend=pend, num_to_optimize=16,
                                        config={"maxiter": 8, "step_limit": 8, "batch_size": 64})
            picks.append(x)
            pend = x if pend is None else np.vstack([pend, x])
        ans, (m, v) = agent.suggest_answers(1, X0, Y0, None, num_starts=32, num_options=512, config={"maxiter": 10})
        return np.vstack(picks), ans, m
    (pg, ag, mg), (pr, ar, mr) = run_both(flow)
    assert np.abs(pg - pr).max() < 1e-5
    assert np.abs(ag - ar).max() < 1e-5 and np.abs(mg - mr).max() < 1e-8


def test_loss_classes_match_oracle_engine():
    def flow():
        agent, X0, Y0 = build(2, 40, "braninhoo", 3, "confidence_bound", lower=False, beta=3.0)
        pools = np.random.RandomState(4).rand(9, 3, 2)
        z = np.random.RandomState(5).randn(64, 3)
        l, (m, c, L) = agent.loss_fn(pools, X0, Y0, fantasy_rvs=z, model=agent.model)
        l1, _ = agent.loss_fn(pools[:, :1], X0, Y0, model=agent.model)
        est = agent.loss_fn._monte_carlo(m, c, base_rvs=z, model=agent.model)
        return l, m, c, L, l1, est
    g, r = run_both(flow)
    for a, b in zip(g, r):
        assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(b).max())


def test_update_fits_hyperparameters_like_the_oracle_engine():
    def flow():
        agent, X0, Y0 = build(2, 60, "braninhoo", 5, "negative_ei")
        Y0 = (Y0 - Y0.mean()) / Y0.std()
        # the default (-inf, 0] bound lets log_noise drift without limit (horseshoe prior); bound it so
        # that both runs stop at the same well-defined optimum
        agent.model.spo_config["var_to_bounds"]["log_noise"] = (np.log(1e-6), 0.0)
        vals = agent.update(None, X0, Y0)
        return np.concatenate([np.ravel(v) for v in vals])
    g, r = run_both(flow)
    assert np.abs(g - r).max() < 1e-4


def test_incremental_greedy_matches_oracle_engine():
    from tests.test_src_mirror import incremental_greedy

    def flow():
        agent, X0, Y0 = build(3, 40, "hartmann3", 7, "confidence_bound", beta=2.0)
        picks, Xf, Yf = incremental_greedy(agent, X0, Y0, 3, 16, num_to_optimize=8, config={"maxiter": 6})
        return picks, Yf
    (pg, yg), (pr, yr) = run_both(flow)
    assert np.abs(pg - pr).max() < 1e-5 and np.abs(yg - yr).max() < 1e-6


def test_successive_halving_and_gd_variants():
    """_optimize_inputs_halving (cached posterior moments + maf_mc_from_moments) and the GradientDescent / Momentum
    branches of _optimize_inputs_sgd, CUDA engine vs the oracle-backed engine on identical host-RNG streams."""
    def flow():
        agent, X0, Y0 = build(3, 40, "hartmann3", 5, "simple_regret")
        agent.loss_fn.num_fantasies = 64
        xh, lh = agent.suggest_inputs(2, X0, Y0, None, num_options=16, subroutine=agent._optimize_inputs_halving,
                                      config={"eval_limit": 128, "min_samples": 16})
        xm, lm = agent.suggest_inputs(2, X0, Y0, None, num_to_optimize=8, subroutine=agent._optimize_inputs_sgd,
                                      config={"step_limit": 6, "batch_size": 32, "method": "train.MomentumOptimizer",
                                              "learning_rate": 0.05})
        return xh, lh, xm, lm
    (xhg, lhg, xmg, lmg), (xhr, lhr, xmr, lmr) = run_both(flow)
    assert np.abs(xhg - xhr).max() < 1e-9 and abs(lhg - lhr) < 1e-8 * max(1.0, abs(lhr))
    assert np.abs(xmg - xmr).max() < 1e-6 and abs(lmg - lmr) < 1e-8 * max(1.0, abs(lmr))


def test_draw_samples_reference_code_golden():
    """models.gaussian_process.draw_samples of the mirror (on the CUDA engine) against golden vectors written by the
    reference's own draw_samples (:328-349; tests/golden/make_golden_acq.py) with the same base samples: at a point
    set, and at one point under fantasy-padded (rank-4) observations as the incremental greedy script calls it."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_acq.npz"))
    r = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("draw/")}
    runtime.set_engine_factory(None)
    runtime.release_engines()
    try:
        model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(r["lenscales"]),
                                        log_amplitude=float(np.log(r["amplitude"])), log_noise=float(np.log(r["noise"])),
                                        mean=float(r["mean"]))
        s2 = model.draw_samples(r["eps"].shape[0], r["X1"], r["X0"], r["Y0"], base_rvs=r["eps"])
        F, n = r["Yf"].shape
        s4 = model.draw_samples(1, r["x1"], np.tile(r["X0"][None, None], [F, 1, 1, 1]), r["Yf"][:, None, :, None],
                                base_rvs=r["eps1"])
    finally:
        runtime.release_engines()
    assert s2.shape == r["samples"].shape and np.abs(s2 - r["samples"]).max() < 1e-9 * np.abs(r["samples"]).max()
    assert np.shape(s4) == r["samples4"].shape and np.abs(s4 - r["samples4"]).max() < 1e-9 * np.abs(r["samples4"]).max()


def test_random_subroutine_on_the_cuda_engine():
    """_optimize_inputs_random (:337-374): forward-only evaluation of uniform shards, the best shard_size sets kept;
    CUDA engine vs the oracle-backed engine on identical host-RNG streams, q > 1 (Monte-Carlo) and q == 1 (closed form)."""
    def flow():
        agent, X0, Y0 = build(3, 48, "hartmann3", 11, "negative_ei")
        agent.loss_fn.num_fantasies = 128
        xq, lq = agent.suggest_inputs(2, X0, Y0, None, num_options=32, subroutine=agent._optimize_inputs_random,
                                      config={"eval_limit": 256})
        x1, l1 = agent.suggest_inputs(1, X0, Y0, None, num_options=64, subroutine=agent._optimize_inputs_random,
                                      config={"eval_limit": 512})
        return xq, lq, x1, l1
    (xqg, lqg, x1g, l1g), (xqr, lqr, x1r, l1r) = run_both(flow)
    assert xqg.shape == (2, 3) and x1g.shape == (1, 3)
    assert np.array_equal(xqg, xqr) and abs(lqg - lqr) < 1e-9 * max(1.0, abs(lqr))
    assert np.array_equal(x1g, x1r) and abs(l1g - l1r) < 1e-9 * max(1.0, abs(l1r))


def _lbfgs_problem(d=3, n=48, seed=21, loss="negative_ei", **kw):
    agent, X0, Y0 = build(d, n, "hartmann3" if d == 3 else "braninhoo", seed, loss, **kw)
    # starts the way the agent draws them: in proportion to the marginal loss (flat-tail starts are rare)
    starts = agent.sample_starts(16, 1, None, X0, Y0, eval_limit=2048)
    return agent, X0, Y0, starts


def test_device_lbfgs_matches_its_restatement_and_beats_the_starts():
    """maf_lbfgs_run (per-restart box-constrained L-BFGS inside the library, no host round trip per evaluation) against
    the test-only NumPy restatement of the same algorithm on the oracle engine: closed form (q = 1), Monte-Carlo with
    fixed base samples and one pending point, and the fantasy-conditioned closed form of the incremental greedy path."""
    def flow():
        agent, X0, Y0, starts = _lbfgs_problem()
        eng = agent.model.bind(X0, Y0)
        prm = agent.loss_fn.acq_params(outputs_old=Y0)
        X1, f1, tr1, info1 = eng.lbfgs_run(prm, starts, None, maxiter=40)
        z = np.random.RandomState(5).randn(64, 2)
        pend = X0[[int(np.argmax(Y0))]] + 0.01
        pools = np.concatenate([np.broadcast_to(pend, (len(starts), 1, 3)), starts], axis=1)
        X2, f2, tr2, info2 = eng.lbfgs_run(prm, pools, z, p_pending=1, maxiter=30)
        F = 8
        Yf = Y0[None, None] + 0.05 * np.random.RandomState(6).randn(F, 1, len(Y0), 1)
        Xf = np.tile(X0[None, None], (F, 1, 1, 1))
        engf = agent.model.bind(Xf, Yf)
        X3, f3, tr3, info3 = engf.lbfgs_run(agent.loss_fn.acq_params(outputs_old=None), starts, None, fantasies=True, maxiter=30)
        return X1, f1, tr1[0], X2, f2, tr2[0], X3, f3, tr3[0]
    g, r = run_both(flow)
    for k in (0, 3, 6):                                     # iterates, final losses, start losses of the three runs
        assert g[k].shape == r[k].shape
        assert np.abs(g[k] - r[k]).max() < 1e-5, k
        assert np.abs(g[k + 1] - r[k + 1]).max() < 1e-7 * max(1.0, np.abs(r[k + 1]).max())
        assert np.abs(g[k + 2] - r[k + 2]).max() < 1e-9 * max(1.0, np.abs(r[k + 2]).max())
        assert np.all(g[k + 1] <= g[k + 2] + 1e-15)         # monotone per restart
        assert g[k + 1].min() < g[k + 2].min()              # and the best restart improved
        assert np.all(g[k] >= 0) and np.all(g[k] <= 1)
    assert np.array_equal(g[3][:, 0], np.broadcast_to(g[3][0, 0], g[3][:, 0].shape))     # pending row untouched


def test_device_lbfgs_is_no_worse_than_host_scipy_on_the_cuda_engine():
    """Same starts, CUDA engine: the device loop against the reference-style joint SciPy L-BFGS-B driving device
    value + gradient (config device=False).  The device loop finds an optimum at least as good."""
    runtime.set_engine_factory(None)
    runtime.release_engines()
    try:
        agent, X0, Y0, starts = _lbfgs_problem(seed=23)
        best = {}
        for device in (True, False):
            handle = agent.appraise_inputs(None, starts, X0, Y0, inputs_as_var=True)[0]
            agent._optimize_inputs_scipy(None, handle, handle, dict(), config={"maxiter": 100, "device": device})
            best[device] = float(np.ravel(handle.evaluate(None)[0]).min())
        assert best[True] <= best[False] + 1e-6 * max(1.0, abs(best[False]))
        ans, (m, v) = agent.suggest_answers(2, X0, Y0, None, num_starts=32, num_options=512, config={"maxiter": 30}, in_order=True)
        assert ans.shape == (2, 3) and m[0] <= m[1] and m[0] <= float(Y0.min()) + 0.1
    finally:
        runtime.release_engines()
