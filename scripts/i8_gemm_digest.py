"""SHA-256 of the INT8 contraction's results on a fixed set of shapes (full / triangular, several digit counts): two
builds of the library that schedule the same integer sums differently must print the same digests."""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200.engine import Engine

e = Engine(0)
for (J, N, ns) in [(256, 256, 6), (512, 640, 6), (256 * 80, 128 * 5, 6), (256 * 160, 2048, 6), (768, 384, 5), (512, 640, 7), (512, 256, 4)]:
    rng = np.random.RandomState(J + N)
    A = rng.rand(J, N) - 0.3
    T0 = rng.randn(N, N)
    for tri, T in ((0, T0), (1, np.tril(T0)), (2, np.triu(T0))):
        C = e.gemm_nt_i8(A, T, tri, ns=ns, lmax=ns + 1)
        err = np.abs(C - A @ T.T).max() / (np.abs(A).max() * np.abs(T).max() * np.sqrt(N))
        print(J, N, ns, tri, hashlib.sha256(C.tobytes()).hexdigest()[:16], "err/scale %.1e" % err, flush=True)
e.close()
