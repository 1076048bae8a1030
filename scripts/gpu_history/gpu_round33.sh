#!/bin/bash
# occupancy / unroll variants of the two covariance kernels (build_variants/libmaf_*.so), then the accuracy report
mkdir -p gpurun_out
L=maximizingacquisitionfunctions_b200/libmaf_b200.so
cp $L /tmp/libmaf_keep.so
for v in A B C D E; do
cp build_variants/libmaf_$v.so $L
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var$v.log 2>&1; echo "variant=$v rc=$?"
tail -1 gpurun_out/bench_var$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'cross_fwd %.2f cross_bwd %.2f' % (k['cross_fwd'], k['cross_bwd']))"
done
cp /tmp/libmaf_keep.so $L
timeout 900 python scripts/check_accuracy.py > gpurun_out/accuracy_n4096.jsonl 2>gpurun_out/accuracy_err.log; echo "accuracy rc=$?"; cat gpurun_out/accuracy_n4096.jsonl | cut -c1-330
