#!/bin/bash
# bench A/B of prebuilt library variants (maximizingacquisitionfunctions_b200/libv_*.so) against the tree's build, one call
mkdir -p gpurun_out/r2
O=gpurun_out/r2
L=maximizingacquisitionfunctions_b200
cp $L/libmaf_b200.so /tmp/libbase.so
for v in base $(ls $L/libv_*.so | sed 's/.*libv_//; s/\.so//') base; do
  if [ $v = base ]; then cp /tmp/libbase.so $L/libmaf_b200.so; else cp $L/libv_$v.so $L/libmaf_b200.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/z_bench_$v.log 2>&1; rc=$?
  tail -1 $O/z_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print('$v', 'rc=$rc', '%.2f ms/step' % d['ms_per_step'], d['clocks']['sm_mhz'], 'cross_fwd %.2f cross_bwd %.2f gemm %.2f+%.2f post %.2f vbar %.2f' % (k['cross_fwd'], k['cross_bwd'], k['gemm_fwd'], k['gemm_bwd'], k['posterior'], k['vbar']))"
done
cp /tmp/libbase.so $L/libmaf_b200.so
