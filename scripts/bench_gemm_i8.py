"""Throughput of the INT8 digit-sliced contraction inside the pipeline shapes (J x 4096 x 4096)."""
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
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    rng = np.random.RandomState(0)
    A = rng.rand(J, N)
    T = np.tril(rng.randn(N, N))
    dA = eng.alloc(A.nbytes).upload(A)
    dT = eng.alloc(T.nbytes).upload(T)
    dC = eng.alloc(A.nbytes)
    ns = int(os.environ.get("OZ_BENCH_NS", "7"))                 # digit planes (levels <= ns + 1)
    for tri, name in ((0, "full"), (1, "lower")):
        eng.profile_enable(True)
        eng.profile_reset()
        for rep in range(3):
            eng._ck(lib.maf_gemm_nt_i8_dev(ctx, J, N, N, dA.ptr, N, dT.ptr, N, dC.ptr, N, tri, ns, ns + 1))
        prof = eng.profile_read()
        eng.profile_enable(False)
        ms = prof["gemm_fwd"][0] / 3
        flops = 2.0 * J * N * N * (1.0 if tri == 0 else 0.5)
        print(json.dumps({"bench": "maf_gemm_nt_i8", "tri": name, "J": J, "N": N, "gemm_ms": ms,
                          "slice_ms": prof["cross_fwd"][0] / 3, "fp64_equiv_tflops_algorithmic": flops / ms * 1e-9,
                          "ns": ns, "int8_tops": flops * (ns * (ns + 1) // 2) / ms * 1e-9 * (1.0 if tri == 0 else (1 + 128.0 / N))}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
