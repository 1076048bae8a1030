"""Objective functions of the BASELINE configs on [0,1]^d (NumPy; data generators only).
TensorFlow-free restatements of the reference's ``tasks/*._numpy`` bodies
(tasks/hartmann3.py:22-61, tasks/hartmann6.py:23-60, tasks/braninhoo.py:22-56,
tasks/levy.py:23-50); checked against reference-generated values in tests/golden/ref_tasks.npz."""
import numpy as np

MINIMA = {"hartmann3": -3.86278, "hartmann6": -3.32237, "braninhoo": 0.397887, "levy": 0.0}

_H_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_H3_A = np.array([[3.0, 10, 30], [0.1, 10, 35], [3.0, 10, 30], [0.1, 10, 35]])
_H3_P = np.array([[3689, 1170, 2673], [4699, 4387, 7470], [1091, 8732, 5547], [381, 5743, 8828]]) * 1e-4
_H6_A = np.array([[10.00, 3.00, 17.00, 3.50, 1.70, 8.00],
                  [0.05, 10.00, 17.00, 0.10, 8.00, 14.00],
                  [3.00, 3.50, 1.70, 10.00, 17.00, 8.00],
                  [17.00, 8.00, 0.05, 10.00, 0.10, 14.00]])
_H6_P = np.array([[1312, 1696, 5569, 124, 8283, 5886],
                  [2329, 4135, 8307, 3736, 1004, 9991],
                  [2348, 1451, 3522, 2883, 3047, 6650],
                  [4047, 8828, 8732, 5743, 1091, 381]]) * 1e-4


def _hartmann(x, A, P):
    y = 0.0
    for i in range(4):
        y = y + _H_ALPHA[i] * np.exp(-np.dot(np.square(x - P[i]), A[i]))
    return -np.expand_dims(y, -1)


def hartmann3(x):
    return _hartmann(np.asarray(x), _H3_A, _H3_P)


def hartmann6(x):
    return _hartmann(np.asarray(x), _H6_A, _H6_P)


def braninhoo(x):
    x = np.asarray(x)
    c1, c2, c3 = -5.1 / (4.0 * np.pi ** 2), 5.0 / np.pi, 10.0 - 10.0 / (8.0 * np.pi)
    z0 = 15 * x[..., [0]]
    z1 = 15 * x[..., [1]] - 5
    return np.square(z1 + c1 * np.square(z0) + c2 * z0 - 6.0) + c3 * np.cos(z0) + 10.0


def levy(x):
    x = np.asarray(x)
    z = 5.0 * x - 1.75
    qq = np.pi * z
    r = np.square(z - 1)
    sin2 = lambda v: np.square(np.sin(v))
    return (sin2(qq[..., :1]) + np.sum(r[..., :-1] * (1 + 10 * sin2(1 + qq[..., :-1])), -1, keepdims=True)
            + r[..., -1:] * (1 + sin2(2 * qq[..., -1:])))
