"""Optimizer-loop latency on the small configurations (launch-bound): maf_adam_steps in blocks of 16 steps with a loss
trace, the way _optimize_inputs_sgd drives it, with the replayed step (CUDA graph, default) and with MAF_GRAPH=0."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200.engine import Engine, acq_params  # noqa: E402
from maximizingacquisitionfunctions_b200.synthetic import hyperparameters, synthetic_inputs  # noqa: E402

CONFIGS = [("C1 n=32 d=3 q=2 S=64 M=128", 32, 3, 2, 64, 128), ("C2 n=256 d=2 q=4 S=1024 M=256", 256, 2, 4, 1024, 256),
           ("C3 n=1024 d=6 q=8 S=4096 M=128", 1024, 6, 8, 4096, 128)]


def main():
    steps, block = 192, 16
    for name, n, d, q, S, M in CONFIGS:
        X0, y0, X, z = synthetic_inputs(n, d, q, S, M, seed=0)
        zs = np.random.RandomState(3).randn(steps, M, q)
        res = {}
        for graph in ("1", "0"):
            os.environ["MAF_GRAPH"] = graph
            eng = Engine(0)
            eng.set_model(**hyperparameters(d))
            eng.set_train(X0, y0)
            prm = acq_params("ei", level=float(np.median(y0)))
            best = math.inf
            for rep in range(3):
                eng.adam_begin(prm, X)
                eng.adam_steps(zs[:block], want_trace=True)              # warm-up block (buffers, capture)
                t0 = time.perf_counter()
                for b in range(block, steps, block):
                    eng.adam_steps(zs[b:b + block], want_trace=True)
                best = min(best, (time.perf_counter() - t0) / (steps - block))
                x = eng.adam_end()
            res[graph] = (best, x)
            if graph == "0":                                               # per-kernel device time of one step, both engines
                for mode in (1, 0):
                    eng.set_gemm_mode(mode)
                    eng.set_train(X0, y0)
                    eng.adam_begin(prm, X)
                    eng.adam_steps(zs[:block])
                    eng.profile_enable(True)
                    eng.profile_reset()
                    eng.adam_steps(zs[:block])
                    prof = eng.profile_read()
                    eng.profile_enable(False)
                    eng.adam_end()
                    print(json.dumps({"config": name, "contraction": "i8" if mode else "f64",
                                      "device_us_per_step": {k: round(v[0] * 1e3 / block, 2) for k, v in prof.items() if v[1]}}), flush=True)
            eng.close()
        assert np.array_equal(res["1"][1], res["0"][1])
        print(json.dumps({"config": name, "us_per_step_replayed": res["1"][0] * 1e6, "us_per_step_host_loop": res["0"][0] * 1e6,
                          "evals_per_s_replayed": S * q * M / res["1"][0], "bitwise_equal_iterates": True}), flush=True)


if __name__ == "__main__":
    main()
