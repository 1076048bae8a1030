#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "float32 or chunking or persistent or large_n or digit_planes" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_new.log; grep -E "^E  " gpurun_out/pytest_new.log | head
timeout 600 python bench.py --steps 3 --warmup 3 --dtype float32 --no-cpu-baseline > gpurun_out/bench_f32.log 2>&1; echo "bench f32 rc=$?"
tail -1 gpurun_out/bench_f32.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks'], d['dtype'])
print(d['roofline']['per_kernel_ms_per_step'], d['roofline']['achieved'])"
