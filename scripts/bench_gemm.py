"""Throughput of the FP64 DMMA contraction kernel in isolation (full / lower / upper T)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200.engine import Engine  # noqa: E402


def main():
    eng = Engine(0)
    lib, ctx = eng._lib, eng._ctx
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    rng = np.random.RandomState(0)
    A = rng.randn(J, N)
    T = rng.randn(N, N)
    dA = eng.alloc(A.nbytes).upload(A)
    dT = eng.alloc(T.nbytes).upload(T)
    dC = eng.alloc(A.nbytes)
    for tri, name in ((0, "full"), (1, "lower"), (2, "upper")):
        flops = 2.0 * J * N * N * (1.0 if tri == 0 else 0.5 * (1 + 128.0 / N))
        for rep in range(2):
            eng._ck(lib.maf_gemm_nt_dev(ctx, J, N, N, dA.ptr, N, dT.ptr, N, dC.ptr, N, tri))
        eng.sync()
        eng.timer_start()
        reps = 3
        for rep in range(reps):
            eng._ck(lib.maf_gemm_nt_dev(ctx, J, N, N, dA.ptr, N, dT.ptr, N, dC.ptr, N, tri))
        ms = eng.timer_stop() / reps
        print(json.dumps({"bench": "maf_gemm_nt", "tri": name, "J": J, "N": N, "ms": ms,
                          "tflops_executed": flops / ms * 1e-9,
                          "tflops_algorithmic": (2.0 * J * N * N * (1.0 if tri == 0 else 0.5)) / ms * 1e-9}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
