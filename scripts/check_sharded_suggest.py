"""agent.suggest_inputs on the CUDA engine under torchrun (NCCL): restarts sharded over the ranks inside the product API,
selection record assembled on the device and gathered with one NCCL all-gather.  Writes (suggestion, loss, index) per call;
run once with 1 process and once with N: the files must agree."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maximizingacquisitionfunctions_b200 import dist as mdist, tasks  # noqa: E402
from maximizingacquisitionfunctions_b200.src import agents, losses, models  # noqa: E402


def main():
    rank, ws, local = mdist.world()
    if ws > 1:
        mdist.init_process_group("nccl")
    d, n = 6, 1024
    rng = np.random.RandomState(0)
    X0 = rng.rand(n, d)
    Y0 = tasks.hartmann6(X0) + math.sqrt(1e-3) * rng.randn(n, 1)
    model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(math.sqrt(d) / 4) * np.ones(d),
                                    log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0, device=local)
    out = []
    for loss, kw, q, S in (("negative_ei", {}, 4, 512), ("confidence_bound", {"beta": 2.0}, 1, 256)):
        loss_fn = getattr(losses, loss)(seed=2, num_fantasies=256, **kw)
        agent = agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=3, timer=time.perf_counter)
        pend = None
        for it in range(2):
            t0 = time.perf_counter()
            x, l = agent.suggest_inputs(q, X0, Y0, None, inputs_pend=pend, num_to_optimize=S,
                                        config={"step_limit": 10, "batch_size": 64, "maxiter": 15})
            out.append({"loss_fn": loss, "q": q, "call": it, "x": x.tolist(), "loss": float(l), "index": int(agent.last_argmin),
                        "wall_s": time.perf_counter() - t0})
            pend = x[:1]
    if rank == 0:
        json.dump(out, open(os.environ.get("SHARD_OUT", "gpurun_out/r2/sharded_suggest") + ".%d.json" % ws, "w"))
        print("world", ws, [(o["loss_fn"], o["index"], round(o["loss"], 12), round(o["wall_s"], 3)) for o in out], flush=True)
    mdist.barrier()


if __name__ == "__main__":
    main()
