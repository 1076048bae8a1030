#!/bin/bash
mkdir -p gpurun_out
MAF_GEMM_I8=1 timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_i8.log 2>&1; echo "pytest i8 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_i8.log | head -30
timeout 600 python bench.py --steps 3 --warmup 3 --gemm i8 --no-cpu-baseline > gpurun_out/bench_i8.log 2>&1; echo "bench i8 rc=$?"
tail -1 gpurun_out/bench_i8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d['roofline'],indent=1))"
