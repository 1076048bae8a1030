"""Where the wall time of agent.suggest_inputs goes at the small BASELINE configs (C1, C2): cProfile of the third call."""
import cProfile
import io
import math
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from maximizingacquisitionfunctions_b200 import runtime, tasks  # noqa: E402
from maximizingacquisitionfunctions_b200.src import agents, losses, models  # noqa: E402

for name, c in bench.AGENT_CONFIGS[:2]:
    rng = np.random.RandomState(0)
    X0 = rng.rand(c["n"], c["d"])
    Y0 = getattr(tasks, c["task"])(X0) + math.sqrt(1e-3) * rng.randn(c["n"], 1)
    model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(math.sqrt(c["d"]) / 4) * np.ones(c["d"]),
                                    log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0, device=0)
    loss_fn = getattr(losses, c["loss"])(seed=2, num_fantasies=1024, **c["kw"])
    agent = agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=3, timer=time.perf_counter)
    cfg = {"step_limit": c["steps"], "batch_size": c["M"]}
    for _ in range(2):
        t0 = time.perf_counter()
        agent.suggest_inputs(c["q"], X0, Y0, None, num_to_optimize=c["S"], subroutine=agent._optimize_inputs_sgd, config=cfg)
        print(name, "wall %.4f s" % (time.perf_counter() - t0))
    pr = cProfile.Profile()
    pr.enable()
    agent.suggest_inputs(c["q"], X0, Y0, None, num_to_optimize=c["S"], subroutine=agent._optimize_inputs_sgd, config=cfg)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:6000])
    runtime.release_engines()
