#!/bin/bash
# A/B of library builds under build_variants/ (compile-time macro experiments) on the bench workload
mkdir -p gpurun_out
LIB=maximizingacquisitionfunctions_b200/libmaf_b200.so
cp $LIB /tmp/libmaf_keep.so
for v in ${VARIANTS:-base A B C D base}; do
  cp build_variants/libmaf_$v.so $LIB
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var$v.log 2>&1; echo "variant $v rc=$?"
  tail -1 gpurun_out/bench_var$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], ' '.join('%s %.2f' % (a, k[a]) for a in k))"
done
cp /tmp/libmaf_keep.so $LIB
