#!/bin/bash
mkdir -p gpurun_out
for f in 3 1 0; do
MAF_OZ_FUSE=$f timeout 600 python bench.py --steps 3 --warmup 3 --gemm i8 --no-cpu-baseline > gpurun_out/bench_i8_f$f.log 2>&1; echo "bench i8 fuse=$f rc=$?"
tail -1 gpurun_out/bench_i8_f$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
