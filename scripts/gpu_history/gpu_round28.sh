#!/bin/bash
mkdir -p gpurun_out
MAF_OZ_FUSE=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ozaki.py -m gpu -q -x --no-header -p no:cacheprovider -k "value_and_grad or pending or pipeline or adam or float32" 2>&1 | tail -3
for f in 3 1; do
MAF_OZ_FUSE=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f$f.log 2>&1; echo "fuse=$f rc=$?"
tail -1 gpurun_out/bench_f$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
