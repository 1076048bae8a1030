#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_diet.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_diet.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
