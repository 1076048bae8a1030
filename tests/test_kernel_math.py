"""CPU: the kernels' hand-derived forward/backward formulas (NumPy mirror,
tests/kernel_math_mirror.py) against the oracle's autograd.  Tolerance 1e-9 relative
(norm-wise), the fp64 parity bar of BASELINE.json."""
import math

import numpy as np
import pytest
import torch

from oracle import maf_oracle as orc
from tests.kernel_math_mirror import Mirror


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def build(n, d, q, S, M, seed, kernel_id):
    gp, X0, y0, X, z = orc.synthetic_problem(n, d, q, S, M, seed=seed, kernel_id=kernel_id)
    ts = orc.prepare_train(gp, X0, y0)
    mir = Mirror(kernel_id, math.sqrt(d) / 4, 1.0, 1e-3, 0.0, 1e-6, X0, y0)
    return ts, mir, X, z


CASES = [
    ("ei", {}, {}),
    ("pi", {"temperature": 0.01}, {"temperature": 0.01}),
    ("pi", {}, {}),
    ("sr", {}, {}),
    ("cb", {"beta": 2.0, "lower": True}, {"beta": 2.0, "lower": True}),
    ("cb", {"beta": 3.0, "lower": False}, {"beta": 3.0, "lower": False}),
]


@pytest.mark.parametrize("kernel_id", ["matern52", "squared_exponential", "matern32"])
@pytest.mark.parametrize("acq,okw,mkw", CASES)
def test_value_grad_formulas(kernel_id, acq, okw, mkw):
    ts, mir, X, z = build(n=96, d=3, q=4, S=6, M=48, seed=2, kernel_id=kernel_id)
    lo, go = orc.value_and_grad(ts, orc.Acq(acq, **okw), X, z)
    lm, gm = mir.value_and_grad(acq, X, z, **mkw)
    assert relerr(lm, lo) < 1e-9
    if np.abs(go).max() > 0:
        assert relerr(gm, go) < 1e-9
    else:
        assert np.abs(gm).max() == 0


def test_pending_weights_incumbents():
    ts, mir, X, z = build(n=64, d=2, q=5, S=4, M=40, seed=4, kernel_id="matern52")
    rng = np.random.RandomState(0)
    w = rng.rand(40)
    w /= w.sum()
    inc = ts.Y0.min().item() + 0.3 * rng.randn(40)
    pend = X[0, :2]
    Xn = X[:, 2:]
    pools = np.concatenate([np.broadcast_to(pend, (4, 2, 2)), Xn], axis=1)
    for acq in ("ei", "sr", "cb"):
        lo, go = orc.value_and_grad(ts, orc.Acq(acq), Xn, z, X_pend=pend, weights=w, incumbents=inc)
        lm, gm = mir.value_and_grad(acq, pools, z, p=2, weights=w, incumbents=inc)
        assert relerr(lm, lo) < 1e-9
        assert relerr(gm, go) < 1e-9


@pytest.mark.parametrize("acq,kw", [("ei", {}), ("pi", {}), ("sr", {}), ("cb", {"lower": True}), ("cb", {"lower": False})])
def test_closed_form_formulas(acq, kw):
    ts, mir, X, _ = build(n=80, d=3, q=1, S=12, M=4, seed=6, kernel_id="matern52")
    lo, go = orc.value_and_grad(ts, orc.Acq(acq, **kw), X, np.zeros((1, 1)))
    lm, gm = mir.closed_form(acq, X, **kw)
    assert relerr(lm, lo) < 1e-9
    assert relerr(gm, go) < 1e-9


def test_duplicate_candidates_split_gradient():
    """reduce_min gradient splits equally among exact ties (two identical pool rows)."""
    ts, mir, X, z = build(n=48, d=2, q=3, S=3, M=32, seed=8, kernel_id="squared_exponential")
    X[:, 1] = X[:, 0]
    lo, go = orc.value_and_grad(ts, orc.Acq("sr"), X, z)
    lm, gm = mir.value_and_grad("sr", X, z)
    assert relerr(lm, lo) < 1e-8
