"""CPU tests of the ``src`` mirror's host logic (agents / losses / models) on a test-only,
oracle-backed engine (tests/fake_engine.py).  The same flows run on the real engine in
tests/test_gpu_mirror.py."""
import math

import numpy as np
import pytest

from maximizingacquisitionfunctions_b200 import runtime
from maximizingacquisitionfunctions_b200.src import agents, losses, models
from oracle import maf_oracle as orc
from tests.fake_engine import FakeEngine


@pytest.fixture(autouse=True)
def fake_engine():
    runtime.set_engine_factory(lambda device, dtype: FakeEngine(device, dtype))
    yield
    runtime.set_engine_factory(None)


def make(d=2, n=24, seed=0, loss="negative_ei", **loss_kwargs):
    rng = np.random.RandomState(seed)
    X0 = rng.rand(n, d)
    Y0 = np.sin(3 * X0.sum(-1, keepdims=True)) + 0.03 * rng.randn(n, 1)
    model = models.gaussian_process(kernel_id="matern52", seed=seed + 1, log_lenscales=np.log(math.sqrt(d) / 4) * np.ones(d),
                                    log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0)
    loss_fn = getattr(losses, loss)(seed=seed + 2, **loss_kwargs)
    agent = agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=seed + 3)
    return agent, X0, Y0


def test_rng_streams_and_prepare_shapes():
    agent, X0, Y0 = make()
    lf = agent.loss_fn
    s0 = int(lf.seed)
    ph, feed = lf.prepare(None, X0, Y0, parallelism=3, num_fantasies=5)
    assert feed["fantasy_rvs"].shape == (5, 3) and int(lf.seed) == s0 + 1
    assert np.array_equal(feed["fantasy_rvs"], np.random.RandomState(s0 + 1).randn(5, 3))
    ph, feed = lf.prepare(None, X0, Y0, parallelism=1)
    assert feed == {} and int(lf.seed) == s0 + 1          # q == 1 draws nothing (myopic_loss.py:307)


def test_loss_return_shapes_match_reference():
    agent, X0, Y0 = make()
    pools = np.random.RandomState(5).rand(7, 3, 2)
    z = np.random.RandomState(6).randn(16, 3)
    losses_, (means, cov, chol) = agent.loss_fn(pools, X0, Y0, fantasy_rvs=z, model=agent.model)
    assert losses_.shape == (7,) and means.shape == (7, 3, 1) and cov.shape == (7, 3, 3) and chol.shape == (7, 3, 3)
    l1, (m1, v1, c1) = agent.loss_fn(pools[:, :1], X0, Y0, model=agent.model)
    assert l1.shape == (7, 1, 1) and m1.shape == (7, 1, 1) and v1.shape == (7, 1, 1) and c1 is None


def test_model_predict_shapes():
    agent, X0, Y0 = make()
    X1 = np.random.RandomState(1).rand(5, 2)
    m, c = agent.model.predict(X1, X0, Y0, full_cov=True)
    assert m.shape == (5, 1) and c.shape == (5, 5)
    m, v = agent.model.predict(X1, X0, Y0)
    assert m.shape == (5, 1) and v.shape == (5, 1)
    m, c = agent.model.predict(X1.reshape(1, 5, 2).repeat(3, 0), X0, Y0, full_cov=True)
    assert m.shape == (3, 5, 1) and c.shape == (3, 5, 5)
    L = agent.model.compute_cholesky(X0, noisy=True)
    assert L.shape == (24, 24) and np.allclose(np.triu(L, 1), 0)


def test_sample_starts_stream_and_shape():
    agent, X0, Y0 = make()
    s0 = int(agent.seed)
    starts = agent.sample_starts(6, 2, None, X0, Y0, eval_limit=1024, shard_size=256)
    assert starts.shape == (6, 2, 2) and starts.min() >= 0 and starts.max() <= 1
    # the k-th shard is RandomState(seed + k).rand(256, d) (module.rng semantics)
    opts = np.vstack([np.random.RandomState(s0 + k).rand(256, 2) for k in range(1, 5)])
    assert all(any(np.array_equal(pt, o) for o in opts) for pt in starts.reshape(-1, 2))


def test_sample_starts_sweeps_stay_powers_of_two():
    """Without a time limit the sweep doubles its shards per device call (64, 128, ... 1024); a remainder that is not a
    power of two is split, so that every sweep stays an exact multiple of the loss's shard size (exact_sharding of the
    reference's sharded_fn, src/misc/sharded_fn.py:56-83) -- 1536 shards used to end in one sweep of 576 x 256 candidates.
    The stream is the reference's whatever the grouping: shard k is RandomState(seed + k).rand(256, d)."""
    agent, X0, Y0 = make()
    s0 = int(agent.seed)
    sizes = []
    handle = agent.appraise_inputs(sess=None, inputs_new=np.empty([0, 1, 2]), inputs_old=X0, outputs_old=Y0)[0]
    evaluate = handle.evaluate

    def spy(feed_dict, inputs_new=None, **kw):
        sizes.append(len(inputs_new))
        return evaluate(feed_dict, inputs_new=inputs_new, **kw)

    handle.evaluate = spy
    starts = agent.sample_starts(5, 2, None, X0, Y0, losses_op=handle, eval_limit=1536 * 256, shard_size=256)
    assert sizes == [g * 256 for g in (64, 128, 256, 512, 512, 64)]
    pts = starts.reshape(-1, 2)
    assert starts.shape == (5, 2, 2) and pts.min() >= 0 and pts.max() <= 1
    opts = np.vstack([np.random.RandomState(s0 + k).rand(256, 2) for k in range(1, 1537)])
    assert all((opts == pt).all(axis=1).any() for pt in pts)


@pytest.mark.parametrize("loss,kw", [("negative_ei", {}), ("negative_pi", {"temperature": 0.01}),
                                     ("simple_regret", {}), ("confidence_bound", {"beta": 2.0})])
def test_suggest_inputs_sgd_improves_and_returns_argmin(loss, kw):
    agent, X0, Y0 = make(loss=loss, **kw)
    agent.loss_fn.num_fantasies = 64
    cfg = {"step_limit": 12, "batch_size": 32}
    X, l = agent.suggest_inputs(2, X0, Y0, None, num_to_optimize=8, subroutine=agent._optimize_inputs_sgd, config=cfg)
    assert X.shape == (2, 2) and X.min() >= 0 and X.max() <= 1 and np.isfinite(l)


def test_sgd_matches_oracle_adam_trajectory():
    agent, X0, Y0 = make()
    starts = np.random.RandomState(9).rand(5, 2, 2)
    s0 = int(agent.loss_fn.seed)
    nodes, loss_seq = agent.optimize_inputs(None, starts, X0, Y0, subroutine=agent._optimize_inputs_sgd,
                                            config={"step_limit": 4, "batch_size": 16})
    # reference stream: appraise_inputs draws one z (num_fantasies), then one per step (batch 16)
    zs = [np.random.RandomState(s0 + 1 + k).randn(16, 2) for k in range(1, 5)]
    gp = orc.GP(log_lenscales=math.log(math.sqrt(2) / 4), log_amplitude=0.0, log_noise=math.log(1e-3), mean=0.0)
    xo, tro = orc.adam_run(orc.prepare_train(gp, X0, Y0), orc.Acq("ei"), starts, zs)
    assert np.allclose(nodes["inputs_new"].value, xo, atol=1e-12)
    assert np.allclose(loss_seq, tro.ravel(), atol=1e-12)


@pytest.mark.parametrize("method,momentum", [("train.GradientDescentOptimizer", 0.0), ("train.MomentumOptimizer", 0.9)])
def test_sgd_gd_and_momentum_variants(method, momentum):
    """config['method'] other than Adam: lr * (global_step + 1)^-0.7 (bayesian_optimization.py:405-412)."""
    agent, X0, Y0 = make()
    starts = np.random.RandomState(9).rand(5, 2, 2)
    s0 = int(agent.loss_fn.seed)
    nodes, loss_seq = agent.optimize_inputs(None, starts, X0, Y0, subroutine=agent._optimize_inputs_sgd,
                                            config={"step_limit": 4, "batch_size": 16, "method": method, "learning_rate": 0.1})
    zs = [np.random.RandomState(s0 + 1 + k).randn(16, 2) for k in range(1, 5)]
    gp = orc.GP(log_lenscales=math.log(math.sqrt(2) / 4), log_amplitude=0.0, log_noise=math.log(1e-3), mean=0.0)
    xo, tro = orc.sgd_run(orc.prepare_train(gp, X0, Y0), orc.Acq("ei"), starts, zs, lr=0.1, momentum=momentum)
    assert np.allclose(nodes["inputs_new"].value, xo, atol=1e-12)
    assert np.allclose(loss_seq, tro.ravel(), atol=1e-12)


def test_successive_halving_rounds_fold_sample_sets():
    """_optimize_inputs_halving (:697-935): round k scores the survivors with the NEW base samples only, re-using the
    cached posterior moments, and folds the estimates weighted by their sample counts -- so a survivor's final
    loss is the plain MC estimate over all base samples drawn so far."""
    agent, X0, Y0 = make(num_fantasies=32)
    drawn = []

    def batch_generator(batch_size):
        z = np.random.RandomState(100 + len(drawn)).randn(batch_size, 2)
        drawn.append(z)
        return {"fantasy_rvs": z}

    starts = np.random.RandomState(4).rand(8, 2, 2)                       # shard of 8 options, q = 2
    nodes, losses_ = agent.optimize_inputs(None, starts, X0, Y0, subroutine=agent._optimize_inputs_halving,
                                           batch_generator=batch_generator,
                                           config={"eval_limit": 64, "min_samples": 8, "expon_step": 1})
    assert [len(z) for z in drawn] == [8, 8, 16]                          # 8 -> 16 -> 32 samples in total
    best = nodes["inputs_new"].value
    assert best.shape == (8, 2, 2) and best.min() >= 0 and best.max() <= 1
    assert len(losses_) == 16                                             # 64 -> 32 -> 16 survivors
    # the 8 selected options carry the 8 smallest folded losses, each equal to the estimate over all 32 samples
    direct = agent.loss_fn(best, X0, Y0, fantasy_rvs=np.vstack(drawn), model=agent.model)[0]
    assert np.allclose(np.sort(direct), np.sort(losses_)[:8], atol=1e-12)
    assert (0, 2, 8) in agent.halving_costs and (0, 2, 32) in agent.halving_costs


def test_successive_halving_closed_form_single_round():
    agent, X0, Y0 = make()
    X, l = agent.suggest_inputs(1, X0, Y0, None, num_options=16, subroutine=agent._optimize_inputs_halving,
                                config={"eval_limit": 64})
    assert X.shape == (1, 2) and np.isfinite(l)


def test_greedy_pending_only_last_row_moves():
    agent, X0, Y0 = make()
    pend = np.random.RandomState(2).rand(2, 2)
    X, l = agent.suggest_inputs(1, X0, Y0, None, inputs_pend=pend, num_to_optimize=4,
                                subroutine=agent._optimize_inputs_sgd, config={"step_limit": 3, "batch_size": 8})
    assert X.shape == (1, 2)


def test_scipy_subroutine_q1_and_joint():
    agent, X0, Y0 = make()
    X, l = agent.suggest_inputs(1, X0, Y0, None, num_to_optimize=6, config={"maxiter": 5})     # q == 1 -> L-BFGS-B
    assert X.shape == (1, 2) and np.isfinite(l)
    agent.loss_fn.num_fantasies = 32
    X, l = agent.suggest_inputs(2, X0, Y0, None, num_to_optimize=4, subroutine=agent._optimize_inputs_scipy,
                                config={"maxiter": 3})
    assert X.shape == (2, 2)


def test_random_subroutine_keeps_top_shard():
    agent, X0, Y0 = make()
    agent.loss_fn.num_fantasies = 32
    X, l = agent.suggest_inputs(2, X0, Y0, None, num_options=16, subroutine=agent._optimize_inputs_random,
                                config={"eval_limit": 64})
    assert X.shape == (2, 2) and np.isfinite(l)


def test_suggest_answers_minimises_posterior_mean():
    agent, X0, Y0 = make()
    ans, (means, var) = agent.suggest_answers(1, X0, Y0, None, num_starts=16, num_options=256, config={"maxiter": 10})
    assert ans.shape == (1, 2) and means.shape == (1,) and var.shape == (1, 1)
    m_all = agent.model.predict(X0, X0, Y0)[0]
    assert means[0] <= m_all.min() + 1e-6
    ans2, _ = agent.suggest_answers(1, X0, Y0, None, risk=1.0, num_starts=8, num_options=128, config={"maxiter": 5})
    assert ans2.shape == (1, 2)


def test_monte_carlo_from_moments_folded_normal():
    """tests/test_qUCB.py:47-53 through the loss class: q=1 MC UCB -> mu + sqrt(beta var)."""
    qucb = losses.confidence_bound(lower=False, beta=2.0, seed=1, model=models.gaussian_process(jitter=0.0, seed=2))
    rvs = np.random.RandomState(3).randn(200000, 1)
    est = qucb._monte_carlo(np.array([[0.3]]), np.array([[2.0]]), base_rvs=rvs)
    assert abs(float(est) - (0.3 + math.sqrt(2.0 * 2.0))) < 1.5e-2


def test_unsupported_paths_raise():
    agent, X0, Y0 = make()
    with pytest.raises(NotImplementedError):
        agent._optimize_inputs_cmaes()
    with pytest.raises(NotImplementedError):
        losses.negative_ei(use_rand_features=True)


def test_fit_objective_and_gradient_match_oracle_autograd():
    """negative_log_posterior (device NLL + host priors) against the oracle's torch autograd of the
    reference objective (gaussian_process.log_likelihood + default priors)."""
    import torch
    agent, X0, Y0 = make(d=3, n=30)
    m = agent.model
    m.state["log_lenscales"] = np.log([0.3, 0.5, 0.7])
    m.state["log_amplitude"], m.state["log_noise"], m.state["mean"] = 0.2, -4.0, 0.1
    val, grads = m.negative_log_posterior(X0, Y0)
    ll = torch.tensor(np.log([0.3, 0.5, 0.7]), requires_grad=True)
    la = torch.tensor(0.2, dtype=orc.DT, requires_grad=True)
    ln = torch.tensor(-4.0, dtype=orc.DT, requires_grad=True)
    mm = torch.tensor(0.1, dtype=orc.DT, requires_grad=True)
    gp = orc.GP(log_lenscales=ll, log_amplitude=la, log_noise=ln, mean=mm, kernel_id="matern52", jitter=1e-6)
    ref = orc.negative_log_posterior(gp, X0, Y0)
    ref.backward()
    assert abs(val - ref.item()) < 1e-9 * abs(ref.item())
    assert np.allclose(grads["log_lenscales"], ll.grad.numpy(), rtol=1e-8, atol=1e-10)
    assert np.allclose(grads["log_amplitude"], la.grad.item(), rtol=1e-8)
    assert np.allclose(grads["log_noise"], ln.grad.item(), rtol=1e-8)
    assert np.allclose(grads["mean"], mm.grad.item(), rtol=1e-8, atol=1e-10)


def test_fit_improves_objective_and_respects_bounds():
    agent, X0, Y0 = make(d=2, n=40)
    m = agent.model
    m.state["log_lenscales"] = np.zeros(2)
    before, _ = m.negative_log_posterior(X0, Y0)
    vals = agent.update(None, X0, Y0, reset=False)
    after, _ = m.negative_log_posterior(X0, Y0)
    assert after < before
    assert np.all(m.log_lenscales <= np.log(2.0) + 1e-12) and np.all(m.log_lenscales >= np.log(1e-2) - 1e-12)
    assert m.log_noise <= 1e-12 and len(vals) == 4


def incremental_greedy(agent, X0, Y0, num_suggest, num_fantasies, **kw):
    """The reference's fantasy-conditioned greedy driver (scripts/run_bayesopt_incremental.py:138-218),
    restated against the mirror's eager API."""
    inputs_old, outputs_old, picks = X0, Y0, []
    for _ in range(num_suggest):
        x, _ = agent.suggest_inputs(1, inputs_old, outputs_old, None, **kw)
        picks.append(np.atleast_1d(np.squeeze(x)))
        if inputs_old.ndim == 2:
            fant = agent.model.draw_samples(num_fantasies, x, inputs_old, outputs_old)           # [F, 1]
            inputs_old = np.tile(np.vstack([inputs_old, x])[None, None], [num_fantasies, 1, 1, 1])
            outputs_old = np.dstack([np.tile(outputs_old[None, None], [num_fantasies, 1, 1, 1]), fant[:, None, None]])
        else:
            fant = agent.model.draw_samples(1, x, inputs_old, outputs_old)                        # [F,1,1,1]
            inputs_old = np.dstack([inputs_old, np.tile(x[None, None], [num_fantasies, 1, 1, 1])])
            outputs_old = np.dstack([outputs_old, fant])
    return np.asarray(picks), inputs_old, outputs_old


@pytest.mark.parametrize("loss,kw", [("negative_ei", {}), ("confidence_bound", {"beta": 2.0})])
def test_incremental_greedy_with_fantasies(loss, kw):
    agent, X0, Y0 = make(d=2, n=20, loss=loss, **kw)
    picks, Xf, Yf = incremental_greedy(agent, X0, Y0, 3, 8, num_to_optimize=6, config={"maxiter": 5})
    assert picks.shape == (3, 2) and Xf.shape == (8, 1, 23, 2) and Yf.shape == (8, 1, 23, 1)
    assert np.array_equal(Xf[0, 0, 20:], picks)
    # rank-4 losses are the mean over fantasies of the per-fantasy closed forms
    cand = np.random.RandomState(0).rand(5, 1, 2)
    l4 = np.ravel(agent.loss_fn(cand, Xf, Yf, model=agent.model)[0])
    l2 = np.mean([np.ravel(agent.loss_fn(cand, Xf[f, 0], Yf[f, 0], model=agent.model)[0]) for f in range(8)], axis=0)
    assert np.allclose(l4, l2, rtol=1e-10, atol=1e-12)
    m4, v4 = agent.model.predict(cand, Xf, Yf)
    assert m4.shape == (8, 5, 1, 1) and v4.shape == (8, 5, 1, 1)


def test_adam_closed_form_q1():
    agent, X0, Y0 = make()
    starts = np.random.RandomState(4).rand(6, 1, 2)
    l0 = np.ravel(agent.loss_fn(starts, X0, Y0, model=agent.model)[0])
    nodes, seq = agent.optimize_inputs(None, starts, X0, Y0, subroutine=agent._optimize_inputs_sgd,
                                       config={"step_limit": 10})
    l1 = np.ravel(agent.loss_fn(nodes["inputs_new"].value, X0, Y0, model=agent.model)[0])
    assert seq.shape == (60,) and l1.mean() < l0.mean()


def test_draw_samples_matches_reference_code_golden(golden_dir):
    """Host logic of models.gaussian_process.draw_samples (rank 2 and fantasy-padded rank 4) against golden vectors
    from the reference's own draw_samples (:328-349) with the same base samples (oracle-backed engine here; the CUDA
    engine is checked by tests/test_gpu_mirror.py::test_draw_samples_reference_code_golden)."""
    import os
    g = np.load(os.path.join(golden_dir, "ref_acq.npz"))
    r = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("draw/")}
    model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(r["lenscales"]),
                                    log_amplitude=float(np.log(r["amplitude"])), log_noise=float(np.log(r["noise"])),
                                    mean=float(r["mean"]))
    s2 = model.draw_samples(r["eps"].shape[0], r["X1"], r["X0"], r["Y0"], base_rvs=r["eps"])
    F, n = r["Yf"].shape
    s4 = model.draw_samples(1, r["x1"], np.tile(r["X0"][None, None], [F, 1, 1, 1]), r["Yf"][:, None, :, None],
                            base_rvs=r["eps1"])
    assert s2.shape == r["samples"].shape and np.abs(s2 - r["samples"]).max() < 1e-11
    assert np.shape(s4) == r["samples4"].shape and np.abs(s4 - r["samples4"]).max() < 1e-11


def test_fantasies_are_rebound_after_the_training_set_was_reloaded():
    """appraise rank-4 -> predict on 2-D data (reloads the training set, the device drops its fantasies) ->
    appraise the same rank-4 set again: the fantasy cache key must not outlive the device state."""
    agent, X0, Y0 = make(d=2, n=12)
    F = 4
    rng = np.random.RandomState(5)
    Xf = np.tile(X0[None, None], (F, 1, 1, 1))
    Yf = Y0[None, None] + 0.1 * rng.randn(F, 1, len(Y0), 1)
    cand = rng.rand(3, 1, 2)
    l_a = np.ravel(agent.loss_fn(cand, Xf, Yf, model=agent.model)[0])
    agent.model.predict(cand[:, 0], X0[:8], Y0[:8])                   # another (X0, y0): set_train, F = 0
    l_b = np.ravel(agent.loss_fn(cand, Xf, Yf, model=agent.model)[0])
    assert np.array_equal(l_a, l_b)
    agent.model.predict(cand[:, 0], X0, Y0)                           # SAME inputs as the fantasy set, 2-D outputs
    l_c = np.ravel(agent.loss_fn(cand, Xf, Yf, model=agent.model)[0])
    assert np.array_equal(l_a, l_c)


def test_device_lbfgs_restatement_is_no_worse_than_joint_scipy():
    """The per-restart L-BFGS that replaces the joint SciPy L-BFGS-B (here: its test-only restatement on the oracle engine)
    finds the optimum the reference-style joint run finds from the same starts, and never leaves a restart worse than
    it started.  (The joint route returns the best S (iterate, restart) pairs of the whole trajectory, several iterates of
    one restart included, so only the best value is comparable.)"""
    agent, X0, Y0 = make(d=2, n=30, seed=4)
    starts = np.random.RandomState(3).rand(12, 1, 2)
    out = {}
    for device in (True, False):
        handle = agent.appraise_inputs(None, starts, X0, Y0, inputs_as_var=True)[0]
        seq = agent._optimize_inputs_scipy(None, handle, handle, dict(), config={"maxiter": 60, "device": device})
        losses = np.ravel(handle.evaluate(None)[0])
        assert handle.value.shape == starts.shape and np.all(handle.value >= 0) and np.all(handle.value <= 1)
        out[device] = (losses, len(seq))
    start_losses = np.ravel(agent.loss_fn(starts, X0, Y0, model=agent.model)[0])
    assert np.all(out[True][0] <= start_losses + 1e-15)            # monotone per restart
    assert out[True][0].min() <= out[False][0].min() + 1e-6        # best optimum found: no worse than the joint run


def test_device_lbfgs_restatement_with_pending_points_and_fixed_samples():
    agent, X0, Y0 = make(d=2, n=20, seed=6, loss="simple_regret")
    pend = X0[[int(np.argmax(Y0))]] + 0.01             # a poor pending point: the new point decides the loss
    starts = np.random.RandomState(2).rand(5, 1, 2)
    handle, feed = agent.appraise_inputs(None, starts, X0, Y0, inputs_pend=pend, inputs_as_var=True,
                                         loss_kwargs={"num_fantasies": 32})[:2]
    before = np.ravel(handle.evaluate(feed)[0])
    agent._optimize_inputs_scipy(None, handle, handle, feed, config={"maxiter": 25})
    after = np.ravel(handle.evaluate(feed)[0])
    assert np.all(after <= before + 1e-12) and after.min() < before.min()
    assert handle.value.shape == (5, 1, 2)


def test_sgd_subroutine_on_fantasy_padded_observations():
    """`--subroutine sgd` with rank-4 observations (the reference accepts it): Adam on the fantasy-averaged closed form."""
    agent, X0, Y0 = make(d=2, n=16, seed=8, loss="confidence_bound", beta=2.0)
    F = 4
    rng = np.random.RandomState(2)
    Xf = np.tile(X0[None, None], (F, 1, 1, 1))
    Yf = Y0[None, None] + 0.05 * rng.randn(F, 1, len(Y0), 1)
    starts = rng.rand(6, 1, 2)
    handle = agent.appraise_inputs(None, starts, Xf, Yf, inputs_as_var=True)[0]
    before = np.ravel(handle.evaluate(None)[0])
    seq = agent._optimize_inputs_sgd(None, handle, handle, config={"step_limit": 15, "learning_rate": 0.02})
    after = np.ravel(handle.evaluate(None)[0])
    assert len(seq) == 15 * 6 and handle.value.shape == (6, 1, 2)
    assert after.mean() < before.mean() and np.all(handle.value >= 0) and np.all(handle.value <= 1)
