#!/bin/bash
# A/B of schedule knobs of the CTA-pair contraction on the bench workload: env assignments per run
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sched.log 2>&1; echo "$* rc=$?"
  tail -1 gpurun_out/bench_sched.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'gemm_fwd %.2f gemm_bwd %.2f' % (k['gemm_fwd'], k['gemm_bwd']))"
}
while read -r line; do [ -n "$line" ] && run $line; done <<CFG
MAF_OZ_GROUPS=1 MAF_OZ_SJ=4
MAF_OZ_GROUPS=2 MAF_OZ_SJ=2
MAF_OZ_GROUPS=2 MAF_OZ_SJ=1
MAF_OZ_GROUPS=1 MAF_OZ_SJ=4
MAF_OZ_GROUPS=2 MAF_OZ_SJ=2
MAF_OZ_GROUPS=3 MAF_OZ_SJ=2
CFG
MAF_OZ_GROUPS=2 MAF_OZ_SJ=2 timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
