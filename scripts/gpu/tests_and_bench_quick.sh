#!/bin/bash
# full GPU suite, then a short bench line (no CPU baseline, no agent timings)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $O/q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/q_pytest_gpu.log; grep -E "^E  " $O/q_pytest_gpu.log | head -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/q_bench.log 2>&1; echo "bench rc=$?"
tail -1 $O/q_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
