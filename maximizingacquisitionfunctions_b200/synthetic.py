"""Seeded synthetic inputs of the BASELINE configs: legacy MT19937
``numpy.random.RandomState`` streams, hyper-parameters fixed as the reference scripts set them
(scripts/run_bayesopt.py:71-74,102-103: matern52, l = sqrt(d)/4, amplitude 1, noise 1e-3,
mean 0; jitter 1e-6 in float64, src/models/gaussian_process.py:42-44)."""
import math

import numpy as np

CONFIGS = {
    # name: n, d, q, S, M, task
    "C1": dict(n=32, d=3, q=2, S=64, M=128, task="hartmann3"),
    "C2": dict(n=256, d=2, q=4, S=1024, M=256, task="braninhoo"),
    "C3": dict(n=1024, d=6, q=8, S=4096, M=128, task="hartmann6"),
    "C4": dict(n=2048, d=6, q=16, S=8192, M=128, task="levy"),
    "C5": dict(n=4096, d=6, q=8, S=65536, M=1024, task=None),
    "headline": dict(n=4096, d=6, q=8, S=65536, M=128, task=None),
}


def hyperparameters(d, kernel_id="matern52"):
    return dict(kernel_id=kernel_id, lenscales=np.full(d, math.sqrt(d) / 4), amplitude=1.0,
                noise=1e-3, mean=0.0, jitter=1e-6)


def synthetic_inputs(n, d, q, S, M, seed=0, task=None):
    """X0=RS(seed).rand(n,d); y0=RS(seed+1).randn(n,1) (or task(X0)+sqrt(1e-3)*randn);
    X=RS(seed+2).rand(S,q,d); z=RS(seed+3).randn(M,q)."""
    RS = np.random.RandomState
    X0 = RS(seed).rand(n, d)
    if task is None:
        y0 = RS(seed + 1).randn(n, 1)
    else:
        from . import tasks
        y0 = getattr(tasks, task)(X0).reshape(n, 1) + math.sqrt(1e-3) * RS(seed + 1).randn(n, 1)
    X = RS(seed + 2).rand(S, q, d)
    z = RS(seed + 3).randn(M, q)
    return X0, y0, X, z
