"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on identical seeded inputs and identical base samples.  Bar (BASELINE.json):
posterior mean/cov and acquisition value/gradient within 1e-9 relative in fp64 (norm-wise:
max-abs error / max-abs value per tensor), identical selected restart."""
import math
import os

import numpy as np
import pytest
import scipy.linalg as spla
import torch

from maximizingacquisitionfunctions_b200.engine import Engine
from oracle import maf_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module", params=["i8", "f64"])
def eng(request):
    """Every parity test runs on both contraction engines: exact INT8 digit slicing on tcgen05 (the default)
    and the FP64 DMMA kernel."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    e = Engine(0)
    e.set_gemm_mode(1 if request.param == "i8" else 0)
    e.test_gemm_mode = 1 if request.param == "i8" else 0
    yield e
    e.close()


def setup_problem(eng, n, d, q, S, M, seed=0, kernel_id="matern52", mean=0.0):
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=seed, kernel_id=kernel_id)
    gp.mean = mean
    ts = orc.prepare_train(gp, X0, y0)
    eng.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, mean, 1e-6)
    eng.set_train(X0, y0)
    return ts, X0, y0, X, z


def prm(kind, **kw):
    from maximizingacquisitionfunctions_b200.engine import acq_params
    return acq_params(kind, **kw)


# ---------------------------------------------------------------- building blocks
@pytest.mark.parametrize("tri", [0, 1, 2])
def test_gemm_fragment_layout_and_triangular_modes(eng, tri):
    rng = np.random.RandomState(tri)
    J, N = 256, 384
    A = rng.randn(J, N)
    T = rng.randn(N, N)
    if tri == 1:
        T = np.tril(T)
    if tri == 2:
        T = np.triu(T)
    C = eng.gemm_nt(A, T, tri)
    ref = A @ T.T
    assert relerr(C, ref) < 1e-13


def test_training_factors(eng):
    ts, X0, y0, X, z = setup_problem(eng, n=200, d=3, q=2, S=4, M=8, seed=1)
    L0, Linv, gamma = eng.factors()
    Lref = ts.chol.numpy()
    assert relerr(L0, Lref) < 1e-12
    assert relerr(Linv @ Lref, np.eye(200)) < 1e-9
    assert relerr(gamma, spla.solve_triangular(Lref, y0.reshape(-1), lower=True)) < 1e-10


# ---------------------------------------------------------------- posterior
def test_posterior_reference_golden(eng, golden_dir):
    """The reference's own test_gp_predict numbers (tests/test_models.py:229-256): SE kernel,
    l=0.1, a=1.2, noise=1e-3, 64 new x 128 old x 4-D.  The 64x64 joint covariance is checked
    through its 16x16 diagonal blocks (pools of q=16) and all marginals (q=1)."""
    g = np.load(os.path.join(golden_dir, "ref_gp_posterior.npz"))
    for key in ("s11_rn2_ro2_mc", "s12_rn2_ro2_mc", "s11_rn2_ro2_memp", "s11_rn3_ro2_mc"):
        X1, X0, Y0 = g[key + "_X1"], g[key + "_X0"], g[key + "_Y0"]
        mean = None if key.endswith("emp") else 0.25
        eng.set_model("squared_exponential", np.full(4, 0.1), 1.2, 1e-3, mean, 1e-6)
        eng.set_train(X0, Y0)
        X1f = X1.reshape(-1, 64, 4)
        mref = g[key + "_mean"].reshape(-1, 64, 1)
        cref = g[key + "_cov"].reshape(-1, 64, 64)
        for b in range(X1f.shape[0]):
            mu, cov = eng.posterior(X1f[b].reshape(4, 16, 4), full_cov=True)
            assert np.abs(mu.reshape(64, 1) - mref[b]).max() < 1e-11
            for k in range(4):
                blk = cref[b][16 * k:16 * k + 16, 16 * k:16 * k + 16]
                assert np.abs(cov[k] - blk).max() < 1e-11
            mu1, var1 = eng.posterior(X1f[b].reshape(64, 1, 4), full_cov=False)
            assert np.abs(mu1.reshape(64, 1) - mref[b]).max() < 1e-11
            assert np.abs(var1.reshape(64) - np.diag(cref[b])).max() < 1e-11


@pytest.mark.parametrize("kernel_id", ["matern52", "squared_exponential", "matern32"])
@pytest.mark.parametrize("n,d,q,S", [(32, 3, 2, 64), (300, 2, 4, 100), (1024, 6, 8, 64), (130, 7, 16, 9), (257, 12, 3, 5)])
def test_posterior_vs_oracle(eng, kernel_id, n, d, q, S):
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, 8, seed=3, kernel_id=kernel_id)
    mu, cov = eng.posterior(X, full_cov=True)
    with torch.no_grad():
        mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
    assert relerr(mu, mo.numpy()) < TOL
    assert relerr(cov, co.numpy()) < TOL
    mu2, var = eng.posterior(X, full_cov=False)
    assert relerr(var[..., 0], np.diagonal(co.numpy(), axis1=-2, axis2=-1)) < TOL


def test_posterior_empirical_mean(eng):
    ts, X0, y0, X, z = setup_problem(eng, 90, 3, 3, 7, 8, seed=5, mean=None)
    mu, cov = eng.posterior(X)
    with torch.no_grad():
        mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
    assert relerr(mu, mo.numpy()) < TOL and relerr(cov, co.numpy()) < TOL


# ---------------------------------------------------------------- acquisition value + gradient
ACQS = [
    ("ei", {}, {}),
    ("pi", {"temperature": 0.01}, {"temperature": 0.01}),
    ("pi", {}, {}),
    ("sr", {}, {}),
    ("cb", {"beta": 2.0, "lower": True}, {"beta": 2.0, "lower": True}),
    ("cb", {"beta": 3.0, "lower": False}, {"beta": 3.0, "lower": False}),
]
SHAPES = [
    # n, d, q, S, M  -- C1 (Hartmann-3 shape), C2 shape at reduced S, C3 shape at reduced S, ragged
    (32, 3, 2, 64, 128),
    (256, 2, 4, 128, 256),
    (1024, 6, 8, 32, 128),
    (100, 5, 7, 11, 100),
    (130, 4, 16, 6, 33),
]


@pytest.mark.parametrize("acq,okw,gkw", ACQS)
@pytest.mark.parametrize("n,d,q,S,M", SHAPES)
def test_value_and_grad_vs_oracle(eng, acq, okw, gkw, n, d, q, S, M):
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=7)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq, **okw), X, z)
    lg, gg = eng.value_and_grad(prm(acq, **gkw), X, z)
    assert relerr(lg, lo) < TOL, ("loss", relerr(lg, lo))
    if np.abs(go).max() > 0:
        assert relerr(gg, go) < TOL, ("grad", relerr(gg, go))
    else:
        assert np.abs(gg).max() == 0
    assert int(np.argmin(lg)) == int(np.argmin(lo))
    idx, val = eng.argmin_last()
    assert idx == int(np.argmin(lo)) and val == lg[idx]
    mu, cov, chol = eng.last_extras(S, q)
    with torch.no_grad():
        _, (mo, co, Lo) = orc.mc_losses(ts, orc.Acq(acq, **okw), torch.as_tensor(X), torch.as_tensor(z), return_extras=True)
    assert relerr(mu, mo.numpy()) < TOL and relerr(cov, co.numpy()) < TOL and relerr(chol, Lo.numpy()) < 1e-8


def test_pending_weights_incumbents(eng):
    n, d, q, S, M = 64, 2, 5, 20, 40
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=4)
    rng = np.random.RandomState(0)
    w = rng.rand(M)
    w /= w.sum()
    inc = float(y0.min()) + 0.3 * rng.randn(M)
    pend, Xn = X[0, :2], X[:, 2:]
    pools = np.ascontiguousarray(np.concatenate([np.broadcast_to(pend, (S, 2, d)), Xn], axis=1))
    for acq in ("ei", "sr", "cb"):
        lo, go = orc.value_and_grad(ts, orc.Acq(acq), Xn, z, X_pend=pend, weights=w, incumbents=inc)
        lg, gg = eng.value_and_grad(prm(acq), pools, z, p_pending=2, weights=w, incumbents=inc)
        assert gg.shape == (S, 3, d)
        assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL


@pytest.mark.parametrize("acq,kw", [("ei", {}), ("pi", {}), ("sr", {}), ("cb", {"lower": True}), ("cb", {"lower": False, "beta": 3.0})])
def test_closed_form_q1(eng, acq, kw):
    ts, X0, y0, X, z = setup_problem(eng, 200, 3, 1, 300, 4, seed=6)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq, **kw), X, np.zeros((1, 1)))
    lg, gg = eng.value_and_grad(prm(acq, **kw), X, None)
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL


def test_value_only_and_explicit_level(eng):
    ts, X0, y0, X, z = setup_problem(eng, 64, 3, 3, 10, 64, seed=9)
    lo, _ = orc.value_and_grad(ts, orc.Acq("ei", level=-0.5), X, z)
    lg, gg = eng.value_and_grad(prm("ei", level=-0.5), X, z, need_grad=False)
    assert gg is None and relerr(lg, lo) < TOL


def test_chunking_is_invisible(eng):
    """restarts are independent (src/misc/sharded_fn.py): results must not depend on the chunking."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M = 160, 3, 3, 200, 64
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=11)
    l1, g1 = eng.value_and_grad(prm("ei"), X, z)
    os.environ["MAF_CHUNK_ROWS"] = "128"
    try:
        e2 = Engine(0)
    finally:
        del os.environ["MAF_CHUNK_ROWS"]
    e2.set_gemm_mode(eng.test_gemm_mode)
    e2.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
    e2.set_train(X0, y0)
    l2, g2 = e2.value_and_grad(prm("ei"), X, z)
    e2.close()
    assert np.array_equal(l1, l2) and np.array_equal(g1, g2)
    perm = np.random.RandomState(1).permutation(S)
    l3, g3 = eng.value_and_grad(prm("ei"), np.ascontiguousarray(X[perm]), z)
    assert np.array_equal(l3, l1[perm]) and np.array_equal(g3, g1[perm])


def test_adam_loop_vs_oracle(eng):
    n, d, q, S, M, steps = 48, 3, 2, 16, 64, 5
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=13)
    zs = np.random.RandomState(99).randn(steps, M, q)
    xo, tro = orc.adam_run(ts, orc.Acq("ei"), X, zs)
    eng.adam_begin(prm("ei"), X)
    trg = eng.adam_steps(zs, want_trace=True)
    xg = eng.adam_end()
    assert relerr(trg, tro) < 1e-7
    assert np.abs(xg - xo).max() < 1e-7
    assert xg.min() >= 0.0 and xg.max() <= 1.0


@pytest.mark.parametrize("method,momentum", [("gd", 0.0), ("momentum", 0.9)])
def test_gd_and_momentum_loops_vs_oracle(eng, method, momentum):
    """GradientDescent / Momentum(0.9) branches of _optimize_inputs_sgd (:405-412), lr (t+1)^-0.7 decay."""
    n, d, q, S, M, steps = 48, 3, 2, 16, 64, 6
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=14)
    zs = np.random.RandomState(98).randn(steps, M, q)
    xo, tro = orc.sgd_run(ts, orc.Acq("sr"), X, zs, lr=0.05, momentum=momentum)
    eng.adam_begin(prm("sr"), X, lr=0.05, method=method, momentum=momentum)
    trg = eng.adam_steps(zs[:2], want_trace=True)
    trg = np.concatenate([trg, eng.adam_steps(zs[2:], want_trace=True)])     # global_step carries over calls
    xg = eng.adam_end()
    assert relerr(trg, tro) < 1e-7
    assert np.abs(xg - xo).max() < 1e-7
    assert xg.min() >= 0.0 and xg.max() <= 1.0


@pytest.mark.parametrize("method", ["adam", "momentum"])
def test_replayed_steps_equal_host_loop_bitwise(eng, method):
    """maf_adam_steps replays one captured step (CUDA graph) when a call has >= 4 steps: same bits as the loop that
    launches every step from the host (MAF_GRAPH=0), across calls of different length, with and without a trace."""
    n, d, q, S, M = 80, 4, 3, 24, 32
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=21)
    zs = np.random.RandomState(5).randn(17, M, q)
    outs = []
    for graph in ("1", "0"):
        os.environ["MAF_GRAPH"] = graph
        try:
            e = Engine(0)
        finally:
            del os.environ["MAF_GRAPH"]
        e.set_gemm_mode(eng.test_gemm_mode)
        e.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
        e.set_train(X0, y0)
        kw = dict(lr=0.05, method=method, momentum=0.9) if method != "adam" else {}
        e.adam_begin(prm("ei"), X, **kw)
        n0 = e.launch_count
        tr = [e.adam_steps(zs[:6], want_trace=True)]
        per_step = (e.launch_count - n0) / 6.0
        tr.append(e.adam_steps(zs[6:8], want_trace=True))           # short call: host loop in both engines
        e.adam_steps(zs[8:12])                                        # no trace: a different captured step
        tr.append(e.adam_steps(zs[12:17], want_trace=True))
        l_last, _ = e.value_and_grad(prm("ei"), e.adam_peek(), z)    # another entry point between the calls
        outs.append((np.concatenate(tr), e.adam_end(), l_last, per_step))
        e.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    assert outs[0][3] == outs[1][3] + 2                              # step_in and step_out kernels of the replayed form


# degenerate and boundary shapes: one training point, one restart, one MC sample, d = 1 and d = MAF_MAX_D, q = MAF_MAX_Q,
# n one below / on / one above the 128-row tile of the contraction
EDGE_SHAPES = [(1, 1, 2, 1, 1), (2, 1, 8, 3, 2), (127, 16, 16, 2, 5), (128, 16, 2, 1, 3), (129, 1, 3, 2, 1), (1, 3, 16, 4, 1)]


@pytest.mark.parametrize("acq", ["sr", "ei", "cb"])
@pytest.mark.parametrize("n,d,q,S,M", EDGE_SHAPES)
def test_edge_shapes_vs_oracle(eng, acq, n, d, q, S, M):
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=3)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq), X, z)
    lg, gg = eng.value_and_grad(prm(acq), X, z)
    assert lg.shape == (S,) and gg.shape == (S, q, d)
    assert relerr(lg, lo) < TOL, ("loss", relerr(lg, lo))
    if np.abs(go).max() > 0:
        assert relerr(gg, go) < TOL, ("grad", relerr(gg, go))
    else:
        assert np.abs(gg).max() == 0
    idx, val = eng.argmin_last()
    assert idx == int(np.argmin(lo)) and val == lg[idx]


@pytest.mark.parametrize("n,d,S", [(1, 1, 1), (1, 16, 7), (127, 2, 3), (129, 16, 1)])
@pytest.mark.parametrize("acq", ["ei", "pi", "sr", "cb"])
def test_closed_form_edge_shapes(eng, acq, n, d, S):
    ts, X0, y0, X, z = setup_problem(eng, n, d, 1, S, 1, seed=5)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq), X, np.zeros((1, 1)))
    lg, gg = eng.value_and_grad(prm(acq), X, None)
    assert lg.shape == (S,) and gg.shape == (S, 1, d)
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL


def test_all_but_one_pending_at_max_pool(eng):
    """q = MAF_MAX_Q with 15 pending points: one free row per restart, gradient [S, 1, d]."""
    n, d, q, S, M = 130, 3, 16, 5, 17
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=8)
    pend, Xn = X[0, :15], X[:, 15:]
    pools = np.ascontiguousarray(np.concatenate([np.broadcast_to(pend, (S, 15, d)), Xn], axis=1))
    for acq in ("cb", "sr"):
        lo, go = orc.value_and_grad(ts, orc.Acq(acq), Xn, z, X_pend=pend)
        lg, gg = eng.value_and_grad(prm(acq), pools, z, p_pending=15)
        assert gg.shape == (S, 1, d)
        assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL


def test_adam_loop_single_restart_single_sample(eng):
    n, d, q, S, M, steps = 5, 1, 2, 1, 1, 6
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=21)
    zs = np.random.RandomState(97).randn(steps, M, q)
    xo, tro = orc.adam_run(ts, orc.Acq("sr"), X, zs)
    eng.adam_begin(prm("sr"), X)
    trg = eng.adam_steps(zs, want_trace=True)
    xg = eng.adam_end()
    assert relerr(trg, tro) < 1e-7 and np.abs(xg - xo).max() < 1e-7


def test_errors_are_loud(eng):
    from maximizingacquisitionfunctions_b200.engine import MafError
    setup_problem(eng, 32, 2, 2, 4, 8)
    with pytest.raises(MafError):
        eng.value_and_grad(prm("ei"), np.zeros((4, 2, 2)), None)          # MC path needs z
    with pytest.raises(MafError):
        eng.value_and_grad(prm("ei"), np.zeros((4, 17, 2)), np.zeros((8, 17)))
    eng.set_model("squared_exponential", np.full(2, 10.0), 1.0, None, 0.0, 0.0)
    with pytest.raises(MafError):                                          # singular K00 (duplicates, no noise/jitter)
        eng.set_train(np.zeros((8, 2)), np.zeros(8))


# ---------------------------------------------------------------- full-size properties
def test_full_size_properties(eng):
    """BASELINE sizes (n=4096, d=6, q=8): oracle on a few restarts + size-independent checks:
    central finite differences of the GPU's own smooth loss (qSR), PSD covariance, Sigma == K11
    minus what the training set explains (<= prior), and restart independence."""
    n, d, q, S, M = 4096, 6, 8, 1024, 128
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=0)
    lg, gg = eng.value_and_grad(prm("ei"), X, z)
    lo, go = orc.value_and_grad(ts, orc.Acq("ei"), X[:8], z)
    assert relerr(lg[:8], lo) < TOL and relerr(gg[:8], go) < TOL
    mu, cov, chol = eng.last_extras(S, q)
    assert np.linalg.eigvalsh(0.5 * (cov + cov.transpose(0, 2, 1))).min() > -1e-9
    assert np.diagonal(cov, axis1=1, axis2=2).max() <= 1.0 + 1e-12
    ls, gs = eng.value_and_grad(prm("sr"), X[:64], z)
    h = 1e-5
    rng = np.random.RandomState(3)
    for _ in range(4):
        s, i, c = rng.randint(64), rng.randint(q), rng.randint(d)
        Xp, Xm = X[:64].copy(), X[:64].copy()
        Xp[s, i, c] += h
        Xm[s, i, c] -= h
        lp, _ = eng.value_and_grad(prm("sr"), Xp, z, need_grad=False)
        lm, _ = eng.value_and_grad(prm("sr"), Xm, z, need_grad=False)
        fd = (lp[s] - lm[s]) / (2 * h)
        assert abs(fd - gs[s, i, c]) < 1e-5 * max(1.0, abs(fd))
        others = np.delete(np.arange(64), s)
        assert np.array_equal(lp[others], ls[others])


@pytest.mark.parametrize("kernel_id", ["matern52", "squared_exponential"])
def test_headline_size_next_to_training_points(eng, kernel_id):
    """n=4096, d=6, q=8, M=128 with half of the candidates 1e-3 away from training points: the posterior variance
    there is ~ the noise level, so Sigma = K11 - V^T V cancels three digits and the reverse contraction three more.
    Still within 1e-9 of the oracle per tensor AND per restart (scripts/check_accuracy.py is the report version;
    profiles/r01_accuracy_n4096.jsonl: 1e-11 .. 2e-10 measured, the FP64 DMMA engine 3e-12 .. 7e-11)."""
    n, d, q, S, M = 4096, 6, 8, 16, 128
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=0, kernel_id=kernel_id)
    X[: S // 2] = np.clip(X0[: (S // 2) * q].reshape(S // 2, q, d) + 1e-3 * np.random.RandomState(7).randn(S // 2, q, d), 0, 1)
    ts = orc.prepare_train(gp, X0, y0)
    eng.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
    eng.set_train(X0, y0)
    for acq in ("cb", "sr"):
        lg, gg = eng.value_and_grad(prm(acq), X, z)
        mu, cov, chol = eng.last_extras(S, q)
        lo, go = orc.value_and_grad(ts, orc.Acq(acq), X, z)
        with torch.no_grad():
            mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
        assert relerr(mu, mo.numpy().reshape(mu.shape)) < TOL and relerr(cov, co.numpy()) < TOL
        assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
        assert max(relerr(gg[s], go[s]) for s in range(S)) < TOL
        assert int(np.argmin(lg)) == int(np.argmin(lo))


@pytest.mark.parametrize("kernel_id", ["matern52", "squared_exponential", "matern32"])
def test_gp_nll_and_gradient_vs_oracle_autograd(eng, kernel_id):
    """maf_gp_nll (gaussian_process.log_likelihood + tf.gradients, :78-106) vs torch autograd."""
    n, d = 150, 4
    rng = np.random.RandomState(2)
    X0, y0 = rng.rand(n, d), rng.randn(n)
    lls, la, ln, mean = np.log([0.3, 0.5, 0.4, 0.8]), 0.1, -5.0, 0.2
    eng.set_model(kernel_id, np.exp(lls), math.exp(la), math.exp(ln), mean, 1e-6)
    eng.set_train(X0, y0)
    nll, g = eng.gp_nll()
    t = [torch.tensor(v, dtype=orc.DT, requires_grad=True) for v in (lls, la, ln, mean)]
    gp = orc.GP(log_lenscales=t[0], log_amplitude=t[1], log_noise=t[2], mean=t[3], kernel_id=kernel_id, jitter=1e-6)
    ref = -orc.log_likelihood_data(gp, torch.as_tensor(X0), torch.as_tensor(y0).reshape(-1, 1))
    ref.backward()
    gref = np.concatenate([t[0].grad.numpy(), [t[1].grad.item(), t[2].grad.item(), t[3].grad.item()]])
    assert abs(nll - ref.item()) < 1e-10 * abs(ref.item())
    assert relerr(g, gref) < 1e-8


@pytest.mark.parametrize("acq,kw", [("ei", {}), ("pi", {}), ("sr", {}), ("cb", {"lower": True}), ("ei", {"level": -0.3})])
def test_fantasy_conditioned_closed_form(eng, acq, kw):
    """rank-4 observation sets: loss = mean over F fantasies of the per-fantasy closed form
    (gaussian_process._predict_nd :195-206 + appraise_inputs :227-229), value and gradient."""
    n, d, S, F = 140, 3, 70, 9
    ts, X0, y0, X, z = setup_problem(eng, n, d, 1, S, 4, seed=21)
    Y = y0.reshape(1, -1) + 0.2 * np.random.RandomState(1).randn(F, n)
    eng.set_fantasies(Y)
    lg, gg = eng.fantasies_value_and_grad(prm(acq, **kw), X[:, 0, :])
    Xt = torch.as_tensor(X).clone().requires_grad_(True)
    lo = orc.fantasy_closed_form_losses(ts.gp, X0, Y, orc.Acq(acq, **kw), Xt)
    lo.sum().backward()
    assert relerr(lg, lo.detach().numpy()) < TOL
    assert relerr(gg, Xt.grad.numpy()[:, 0, :]) < TOL
    mu, var = eng.fantasies_posterior(X[:, 0, :])
    with torch.no_grad():
        m3, v3 = orc.posterior(orc.prepare_train(ts.gp, X0, Y[3]), torch.as_tensor(X), full_cov=False)
    assert relerr(mu[3], m3.numpy().reshape(-1)) < TOL and relerr(var, v3.numpy().reshape(-1)) < TOL


def test_adam_closed_form_q1_vs_oracle(eng):
    ts, X0, y0, X, z = setup_problem(eng, 60, 3, 1, 12, 4, seed=17)
    xo, tro = orc.adam_run(ts, orc.Acq("ei"), X, [np.zeros((1, 1))] * 4)
    eng.adam_begin(prm("ei"), X)
    trg = eng.adam_steps(None, want_trace=True, steps=4)
    xg = eng.adam_end()
    assert relerr(trg, tro) < 1e-8 and np.abs(xg - xo).max() < 1e-8


def _reference_acq_case_names():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_acq.npz"))
    return [str(c) for c in g["cases"]]


@pytest.mark.parametrize("name", _reference_acq_case_names())
def test_reference_code_golden(eng, name):
    """The CUDA path against golden vectors written by the reference's OWN src.losses / src.models code (run on the
    eager TF shim, tests/golden/make_golden_acq.py): value, posterior moments, Cholesky factor and gradient, 1e-9."""
    from tests.test_oracle import load_reference_acq_case
    rec, gp, acq = load_reference_acq_case(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"), name)
    pools = rec["pools"]
    S, q, d = pools.shape
    eng.set_model(gp.kernel_id, rec["lenscales"], float(rec["amplitude"]), float(rec["noise"]), float(rec["mean"]),
                  float(rec["jitter"]))
    eng.set_train(rec["X0"], rec["Y0"])
    p = prm(acq.kind, level=acq.level, temperature=acq.temperature, beta=acq.beta, lower=acq.lower)
    lg, gg = eng.value_and_grad(p, pools, rec.get("z"), weights=rec.get("weights"), incumbents=rec.get("incumbents"))
    assert relerr(lg, rec["losses"].reshape(-1)) < TOL
    if np.abs(rec["grad"]).max() > 0:
        assert relerr(gg, rec["grad"]) < TOL
    else:
        assert np.abs(gg).max() == 0.0
    if q > 1:
        mu, cov, chol = eng.last_extras(S, q)
        assert relerr(mu, rec["means"]) < TOL and relerr(cov, rec["cov"]) < TOL and relerr(chol, rec["chol"]) < TOL
    else:
        mu, var = eng.posterior(pools, full_cov=False)
        assert relerr(mu, rec["means"]) < TOL and relerr(var, rec["cov"]) < TOL


def _fantasy_case_names():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_acq.npz"))
    return [str(c) for c in g["fantasy_cases"]]


@pytest.mark.parametrize("name", _fantasy_case_names())
def test_fantasy_padded_reference_code_golden(eng, name):
    """maf_fantasies_value_grad against the reference's own rank-4 path (golden vectors, fan_* cases)."""
    from tests.test_oracle import fantasy_case
    rec, gp, acq = fantasy_case(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"), name)
    eng.set_model(gp.kernel_id, rec["lenscales"], float(rec["amplitude"]), float(rec["noise"]), float(rec["mean"]),
                  float(rec["jitter"]))
    eng.set_train(rec["X0"], rec["Y"][0])
    eng.set_fantasies(rec["Y"])
    lg, gg = eng.fantasies_value_and_grad(prm(acq.kind, beta=acq.beta, lower=acq.lower), rec["pools"][:, 0, :])
    assert relerr(lg, rec["losses"].reshape(-1)) < TOL
    assert relerr(gg, rec["grad"][:, 0, :]) < TOL


@pytest.mark.parametrize("name", ["pend_negative_ei", "pend_confidence_bound"])
def test_pending_points_reference_code_golden(eng, name):
    """p_pending > 0 through the C ABI against the reference's own pools construction and loss code (golden vectors)."""
    from tests.test_oracle import pending_case
    rec, gp, acq = pending_case(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"), name)
    eng.set_model(gp.kernel_id, rec["lenscales"], float(rec["amplitude"]), float(rec["noise"]), float(rec["mean"]),
                  float(rec["jitter"]))
    eng.set_train(rec["X0"], rec["Y0"])
    S, p = rec["Xnew"].shape[0], rec["Xpend"].shape[0]
    pools = np.concatenate([np.tile(rec["Xpend"][None], [S, 1, 1]), rec["Xnew"]], axis=1)
    lg, gg = eng.value_and_grad(prm(acq.kind, level=acq.level, beta=acq.beta, lower=acq.lower), pools, rec["z"], p_pending=p)
    assert relerr(lg, rec["losses"]) < TOL
    assert gg.shape == rec["grad"].shape and relerr(gg, rec["grad"]) < TOL


# ---------------------------------------------------------------- the configurations bench.py / run_configs.py time
def per_restart_err(a, b):
    a, b = np.asarray(a), np.asarray(b).reshape(np.shape(a))
    return max(float(np.abs(a[s] - b[s]).max() / max(np.abs(b[s]).max(), 1e-300)) for s in range(a.shape[0]))


def test_bench_workload_qei_median_level(eng):
    """EXACTLY what bench.py times: qEI, n=4096, d=6, q=8, M=128, white-noise outputs, level = median(y0) (every restart
    has a non-zero loss and gradient), uniform random candidates.  64 restarts against the oracle: 1e-9 per tensor AND
    per restart (negative_ei.py:31-74 with an explicit level), identical selected restart."""
    n, d, q, S, M = 4096, 6, 8, 64, 128
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=0)
    level = float(np.median(y0))
    lo, go = orc.value_and_grad(ts, orc.Acq("ei", level=level), X, z)
    lg, gg = eng.value_and_grad(prm("ei", level=level), X, z)
    assert np.abs(lo).min() > 0 and min(np.abs(go[s]).max() for s in range(S)) > 0, "the workload must not be degenerate"
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    assert per_restart_err(lg.reshape(S, 1), lo.reshape(S, 1)) < TOL
    assert per_restart_err(gg, go) < TOL
    mu, cov, chol = eng.last_extras(S, q)
    with torch.no_grad():
        mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
    assert relerr(mu, mo.numpy()) < TOL and relerr(cov, co.numpy()) < TOL
    assert per_restart_err(mu, mo.numpy()) < TOL and per_restart_err(cov, co.numpy()) < TOL
    assert int(np.argmin(lg)) == int(np.argmin(lo))
    idx, _ = eng.argmin_last()
    assert idx == int(np.argmin(lo))


@pytest.mark.parametrize("p", [0, 15])
def test_c4_shape_levy_q16_pending(eng, p):
    """BASELINE config 4: Levy d=6, n=2048, q=16, qLCB (beta=2), M=128; joint (p=0) and the last step of the pending-point
    greedy construction (15 pending points shared by every restart, one trainable row;
    bayesian_optimization.py:1106-1123, scripts/run_bayesopt.py:137-176)."""
    from maximizingacquisitionfunctions_b200 import tasks
    n, d, q, S, M = 2048, 6, 16, 32, 128
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=2, task=tasks.levy)
    ts = orc.prepare_train(gp, X0, y0)
    eng.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
    eng.set_train(X0, y0)
    if p:
        pend, Xn = X[0, :p], X[:, p:]
        pools = np.ascontiguousarray(np.concatenate([np.broadcast_to(pend, (S, p, d)), Xn], axis=1))
        lo, go = orc.value_and_grad(ts, orc.Acq("cb", beta=2.0, lower=True), Xn, z, X_pend=pend)
    else:
        pools = X
        lo, go = orc.value_and_grad(ts, orc.Acq("cb", beta=2.0, lower=True), X, z)
    lg, gg = eng.value_and_grad(prm("cb", beta=2.0, lower=True), pools, z, p_pending=p)
    assert gg.shape == (S, q - p, d)
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    assert per_restart_err(lg.reshape(S, 1), lo.reshape(S, 1)) < TOL and per_restart_err(gg, go) < TOL
    assert int(np.argmin(lg)) == int(np.argmin(lo))


def test_c5_shape_m1024_nondegenerate_level(eng):
    """BASELINE config 5: n=4096, d=6, q=8, M=1024 base samples; level = median(y0) so that no restart is identically 0."""
    n, d, q, S, M = 4096, 6, 8, 32, 1024
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=1)
    level = float(np.median(y0))
    lo, go = orc.value_and_grad(ts, orc.Acq("ei", level=level), X, z)
    lg, gg = eng.value_and_grad(prm("ei", level=level), X, z)
    assert np.abs(lo).min() > 0
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    assert per_restart_err(lg.reshape(S, 1), lo.reshape(S, 1)) < TOL and per_restart_err(gg, go) < TOL
    assert int(np.argmin(lg)) == int(np.argmin(lo))


@pytest.mark.parametrize("task_name,n,d,q,acq,kw", [
    ("hartmann3", 32, 3, 2, "ei", {}),
    ("braninhoo", 256, 2, 4, "pi", {"temperature": 0.01}),
    ("braninhoo", 256, 2, 4, "sr", {}),
    ("hartmann6", 1024, 6, 8, "ei", {}),
])
def test_baseline_configs_on_their_task_outputs(eng, task_name, n, d, q, acq, kw):
    """C1-C3 with the outputs of the task each is named after (tasks.py restates tasks/*._numpy) + sqrt(1e-3) noise, the
    reference's default level min(y0)."""
    from maximizingacquisitionfunctions_b200 import tasks
    S, M = 48, 128
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=3, task=getattr(tasks, task_name))
    ts = orc.prepare_train(gp, X0, y0)
    eng.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
    eng.set_train(X0, y0)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq, **kw), X, z)
    lg, gg = eng.value_and_grad(prm(acq, **kw), X, z)
    assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    assert int(np.argmin(lg)) == int(np.argmin(lo))


def test_digit_switch_api(eng):
    """maf_set_digits: 7/7 is the default grade; fewer planes change the result only inside their documented error."""
    if not eng.test_gemm_mode:
        pytest.skip("INT8 engine only")
    ts, X0, y0, X, z = setup_problem(eng, 512, 4, 4, 16, 64, seed=5)
    l7, g7 = eng.value_and_grad(prm("sr"), X, z)
    eng.set_digits(5, 5)
    l5, g5 = eng.value_and_grad(prm("sr"), X, z)
    eng.set_digits(7, 7)
    l7b, g7b = eng.value_and_grad(prm("sr"), X, z)
    assert np.array_equal(l7, l7b) and np.array_equal(g7, g7b)
    assert 0 < relerr(l5, l7) < 1e-5
    from maximizingacquisitionfunctions_b200.engine import MafError
    with pytest.raises(MafError):
        eng.set_digits(8, 7)


# ---------------------------------------------------------------- level switch of the contractions
def _near_problem(n, d, q, S, M, kernel_id="matern52", seed=0):
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=seed, kernel_id=kernel_id)
    X[: S // 2] = np.clip(X0[: (S // 2) * q].reshape(S // 2, q, d) + 1e-3 * np.random.RandomState(7).randn(S // 2, q, d), 0, 1)
    return gp, X0, y0, X, z


def test_level_switch_keeps_the_bar_per_restart():
    """Default engine (level switch on) at the headline size, half of the candidates 1e-3 from training points: the
    restarts the cheaper level cannot serve are recomputed on all planes inside the call (statistics say so) and value,
    gradient, mean and covariance stay within 1e-9 of the oracle per tensor AND per restart."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M = 4096, 6, 8, 64, 128
    gp, X0, y0, X, z = _near_problem(n, d, q, S, M)
    ts = orc.prepare_train(gp, X0, y0)
    level = float(np.median(y0))
    e = Engine(0)
    try:
        e.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
        e.set_train(X0, y0)
        st0 = e.level_switch_stats(reset=True)
        assert st0["err_cov_abs"] > 0 and st0["err_grad_per_scale"] > 0 and st0["tol"] == 2.5e-10
        for okw, gkw in ((("ei",), {"level": level}), (("cb",), {}), (("sr",), {})):
            lo, go = orc.value_and_grad(ts, orc.Acq(okw[0], **gkw), X, z)
            lg, gg = e.value_and_grad(prm(okw[0], **gkw), X, z)
            mu, cov, _ = e.last_extras(S, q)
            with torch.no_grad():
                mo, co = orc.posterior(ts, torch.as_tensor(X), full_cov=True)
            assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL and relerr(mu, mo.numpy()) < TOL and relerr(cov, co.numpy()) < TOL
            assert per_restart_err(gg, go) < TOL and per_restart_err(cov, co.numpy()) < TOL and per_restart_err(mu, mo.numpy()) < TOL
            assert int(np.argmin(lg)) == int(np.argmin(lo))
        st = e.level_switch_stats()
        assert st["tested_forward"] == 3 * S and st["recomputed_forward"] >= 3 * (S // 2) - 3     # the near half at least
        assert st["recomputed_forward"] <= st["tested_forward"] and st["recomputed_reverse"] <= st["tested_reverse"]
    finally:
        e.close()


def test_level_switch_is_invisible_to_chunking_and_can_be_switched_off():
    """Results with the switch on do not depend on how the restarts are chunked (a restart's result depends on that restart
    only), and agree with the switch-off results to the tolerance the switch promises."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M = 1024, 4, 3, 300, 64
    gp, X0, y0, X, z = _near_problem(n, d, q, S, M, seed=4)
    out = {}
    for tag, env, on in (("a", None, True), ("b", "256", True), ("off", None, False)):
        if env:
            os.environ["MAF_CHUNK_ROWS"] = env
        try:
            e = Engine(0)
            e.set_level_switch(on)
            e.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
            e.set_train(X0, y0)
            out[tag] = e.value_and_grad(prm("cb"), X, z) + e.last_extras(S, q)[:2] + (e.level_switch_stats(),)
            e.close()
        finally:
            os.environ.pop("MAF_CHUNK_ROWS", None)
    for u in range(4):
        assert np.array_equal(out["a"][u], out["b"][u])
    assert out["a"][4]["tested_forward"] == S and out["off"][4]["tested_forward"] == 0
    assert relerr(out["a"][0], out["off"][0]) < 1e-9 and relerr(out["a"][1], out["off"][1]) < 1e-9
    assert per_restart_err(out["a"][1], out["off"][1]) < 1e-9 and per_restart_err(out["a"][3], out["off"][3]) < 1e-9


@pytest.mark.parametrize("kernel_id", ["matern52", "matern32", "squared_exponential"])
def test_stored_kernel_derivatives_route_equals_the_recomputing_kernel(kernel_id):
    """The forward covariance kernel leaves d kappa / d r2 of every pair behind and the reverse kernel streams it
    (cross_bwd_g_kernel, default) instead of recomputing square root and exponential per pair (MAF_G_ROUTE=0): same
    arithmetic for the derivative, so values are identical and gradients agree to rounding; both within the bar of
    the oracle.  Odd n (masked last column of the two-column loop), shared pending rows, per-restart pending rows, a
    candidate ON a training point (gradient blocked by the clamp), and the second pass of the reverse level switch."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M = 1153, 5, 5, 96, 64
    gp, X0, y0, X, z = _near_problem(n, d, q, S, M, kernel_id=kernel_id, seed=11)
    X[3, 1] = X0[17]                                          # r2 = 0 against one training point
    ts = orc.prepare_train(gp, X0, y0)
    Xp = X.copy()
    Xp[:, :2] = X[0, :2]                                      # two shared pending rows
    out = {}
    for route in ("1", "0"):
        os.environ["MAF_G_ROUTE"] = route
        try:
            e = Engine(0)
            e.set_model(kernel_id, np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
            e.set_train(X0, y0)
            out[route] = (e.value_and_grad(prm("cb"), X, z), e.value_and_grad(prm("ei", level=float(np.median(y0))), Xp, z, p_pending=2),
                          e.value_and_grad(prm("sr"), X, z, p_pending=1), e.level_switch_stats())
            e.close()
        finally:
            os.environ.pop("MAF_G_ROUTE", None)
    for u in range(3):
        assert np.array_equal(out["1"][u][0], out["0"][u][0])                     # losses never see the reverse kernel
        # different summation order over the training points (two columns per lane): rounding of cancelling sums only
        assert relerr(out["1"][u][1], out["0"][u][1]) < 1e-11 and per_restart_err(out["1"][u][1], out["0"][u][1]) < 1e-10
    lo, go = orc.value_and_grad(ts, orc.Acq("cb"), X, z)
    assert relerr(out["1"][0][0], lo) < TOL and relerr(out["1"][0][1], go) < TOL and per_restart_err(out["1"][0][1], go) < TOL
    lo, go = orc.value_and_grad(ts, orc.Acq("ei", level=float(np.median(y0))), Xp[:, 2:], z, X_pend=Xp[0, :2])
    assert relerr(out["1"][1][0], lo) < TOL and relerr(out["1"][1][1], go) < TOL
    assert out["1"][2][1].shape == (S, q - 1, d)              # pools with their own pending row: trainable rows only


def test_level_switch_inside_the_replayed_adam_loop():
    """The switch decides on the device, so a captured optimizer step replays it: the replayed loop equals the host-driven
    loop bit for bit with the switch on."""
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M, T = 1024, 6, 2, 40, 32, 6
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=9)
    zs = np.random.RandomState(3).randn(T, M, q)
    res = []
    for graph in ("1", "0"):
        os.environ["MAF_GRAPH"] = graph
        try:
            e = Engine(0)
            e.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
            e.set_train(X0, y0)
            e.adam_begin(prm("sr"), X)
            tr = e.adam_steps(zs, want_trace=True)
            res.append((e.adam_end(), tr, e.level_switch_stats()))
            e.close()
        finally:
            os.environ.pop("MAF_GRAPH", None)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for key in ("tested_forward", "recomputed_forward", "tested_reverse", "recomputed_reverse"):
        assert res[0][2][key] == res[1][2][key]
    assert res[0][2]["tested_forward"] == T * S and res[0][2]["tested_reverse"] == T * S     # the switch was live in both


def test_pending_rows_shared_block_and_per_restart_fallback(eng):
    """Pending points shared by all restarts are contracted once (the pools of the agent: concat(tile(Xpend), new));
    pools whose first p rows differ from restart to restart are legal through the C ABI too and take the per-restart
    route.  Both against the oracle; the shared route must also agree with the per-restart route on the same pools."""
    n, d, q, p, S, M = 300, 3, 5, 2, 40, 48
    ts, X0, y0, X, z = setup_problem(eng, n, d, q, S, M, seed=12)
    shared = X.copy()
    shared[:, :p] = X[0, :p]
    lo, go = orc.value_and_grad(ts, orc.Acq("cb"), shared[:, p:], z, X_pend=shared[0, :p])
    lg, gg = eng.value_and_grad(prm("cb"), shared, z, p_pending=p)
    assert gg.shape == (S, q - p, d) and relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    mu, cov, chol = eng.last_extras(S, q)
    with torch.no_grad():
        mo, co = orc.posterior(ts, torch.as_tensor(shared), full_cov=True)
    assert relerr(mu, mo.numpy()) < TOL and relerr(cov, co.numpy()) < TOL
    # different "pending" rows per restart: gradient of the trainable rows of the full-pool problem
    lo2, go2 = orc.value_and_grad(ts, orc.Acq("cb"), X, z)
    lg2, gg2 = eng.value_and_grad(prm("cb"), X, z, p_pending=p)
    assert relerr(lg2, lo2) < TOL and relerr(gg2, go2[:, p:]) < TOL
    # one restart of the shared pools evaluated alone (S = 1: trivially shared) equals its row in the batch
    l1, g1 = eng.value_and_grad(prm("cb"), shared[7:8], z, p_pending=p)
    assert relerr(l1, lg[7:8]) < 1e-12 and relerr(g1, gg[7:8]) < 1e-10


def test_two_engines_on_two_devices_in_one_process():
    """cudaFuncSetAttribute opt-ins (dynamic shared memory of the factorisation, the Monte-Carlo kernel at q = 16, the
    contraction kernels) are per device: a second engine on another GPU of the same process must work like the first."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from maximizingacquisitionfunctions_b200.engine import Engine
    n, d, q, S, M = 1100, 3, 16, 12, 32
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=6)
    ts = orc.prepare_train(gp, X0, y0)
    lo, go = orc.value_and_grad(ts, orc.Acq("sr"), X, z)
    engines = [Engine(0), Engine(1)]
    try:
        for e in engines:
            e.set_model("matern52", np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-6)
            e.set_train(X0, y0)
        for e in reversed(engines):
            lg, gg = e.value_and_grad(prm("sr"), X, z)
            assert relerr(lg, lo) < TOL and relerr(gg, go) < TOL
    finally:
        for e in engines:
            e.close()
