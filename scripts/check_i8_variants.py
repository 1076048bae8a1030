"""Correctness of an INT8 GEMM variant (MAF_OZ_VARIANT) against NumPy and against variant 3, on sizes with several
tiles per CTA / CTA pair and odd k-chunk counts."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200.engine import Engine

def run(variant, A, T, tri, ns, lmax):
    os.environ["MAF_OZ_VARIANT"] = str(variant)
    e = Engine(0)
    try:
        return e.gemm_nt_i8(A, T, tri, ns=ns, lmax=lmax)
    finally:
        e.close()

v = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ok = True
for (J, N, ns, lmax) in [(256, 128, 7, 8), (512, 640, 7, 8), (256 * 80, 128 * 5, 7, 8), (768, 384, 5, 6), (512, 256, 4, 5), (512, 384, 6, 7), (256, 256, 3, 4)]:
    rng = np.random.RandomState(J + N)
    A = rng.rand(J, N)
    T0 = rng.randn(N, N)
    for tri, T in ((0, T0), (1, np.tril(T0)), (2, np.triu(T0))):
        C = run(v, A, T, tri, ns, lmax)
        C3 = run(3, A, T, tri, ns, lmax)
        ref = A @ T.T
        scale = np.abs(A).max(1, keepdims=True) * np.abs(T).max(1, keepdims=True).T * math.sqrt(N)
        err = (np.abs(C - ref) / scale).max()
        same = np.array_equal(C, C3)
        print("J=%d N=%d ns=%d tri=%d  err/scale=%.2e  identical_to_v3=%s" % (J, N, ns, tri, err, same), flush=True)
        ok &= same
print("ALL IDENTICAL" if ok else "MISMATCH")
