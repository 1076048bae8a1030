#!/bin/bash
# fast sqrt/exp in the covariance kernels: parity, accuracy report at n=4096, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -5
timeout 600 python scripts/check_accuracy.py > gpurun_out/accuracy_n4096.jsonl 2>gpurun_out/accuracy_err.log; echo "accuracy rc=$?"; cat gpurun_out/accuracy_n4096.jsonl | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fastmath.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_fastmath.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
