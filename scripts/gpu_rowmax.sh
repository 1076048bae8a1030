#!/bin/bash
# row maxima of V^T from the forward contraction's epilogue: full GPU suite, bitwise A/B, bench A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import os, numpy as np
from maximizingacquisitionfunctions_b200.engine import Engine, acq_params
from maximizingacquisitionfunctions_b200.synthetic import hyperparameters, synthetic_inputs
n, d, q, S, M = 1000, 6, 8, 300, 128
X0, y0, X, z = synthetic_inputs(n, d, q, S, M, seed=3)
out = []
for rm in ("1", "0"):
    os.environ["MAF_OZ_ROWMAX"] = rm
    e = Engine(0); e.set_model(**hyperparameters(d)); e.set_train(X0, y0)
    out.append(e.value_and_grad(acq_params("ei", level=float(np.median(y0))), X, z)); e.close()
print("bitwise equal with / without epilogue row maxima:", np.array_equal(out[0][0], out[1][0]), np.array_equal(out[0][1], out[1][1]), "max |grad|", np.abs(out[0][1]).max())
PY
run() {
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sched.log 2>&1; echo "$* rc=$?"
  tail -1 gpurun_out/bench_sched.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], ' '.join('%s %.2f' % (a, k[a]) for a in k))"
}
run MAF_OZ_ROWMAX=0
run MAF_OZ_ROWMAX=1
run MAF_OZ_ROWMAX=0
run MAF_OZ_ROWMAX=1
