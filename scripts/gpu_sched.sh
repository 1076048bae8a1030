#!/bin/bash
# A/B of the balanced static tile schedule of the CTA-pair contraction (MAF_OZ_SCHED) on the bench workload
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sched.log 2>&1; echo "$* rc=$?"
  tail -1 gpurun_out/bench_sched.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'gemm_fwd %.2f gemm_bwd %.2f' % (k['gemm_fwd'], k['gemm_bwd']))"
}
run MAF_OZ_SCHED=0
run MAF_OZ_SCHED=1
run MAF_OZ_SCHED=0
run MAF_OZ_SCHED=1 MAF_OZ_SCHED_OVH=0.5
run MAF_OZ_SCHED=1 MAF_OZ_SCHED_OVH=3
run MAF_OZ_SCHED=1 MAF_OZ_SJ=8
MAF_OZ_SCHED=1 timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
