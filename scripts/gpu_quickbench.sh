#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_q.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks'])
print(d['roofline']['per_kernel_ms_per_step'])"
