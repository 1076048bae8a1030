#!/bin/bash
# tile-order knobs of the contraction re-swept on the six-plane first pass (round-1 sweeps were on seven planes): column groups, super-row height
mkdir -p gpurun_out/r2
O=gpurun_out/r2
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/s_bench_$name.log 2>&1; rc=$?
  tail -1 $O/s_bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print('$name rc=$rc %.2f ms/step %.4e sm %.0f gemm %.2f + %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], k['gemm_fwd'], k['gemm_bwd']))"
}
run base_a MAF_OZ_GROUPS=2
run g1 MAF_OZ_GROUPS=1
run g3 MAF_OZ_GROUPS=3
run sj1 MAF_OZ_SJ=1
run sj4 MAF_OZ_SJ=4
run g1sj4 MAF_OZ_GROUPS=1 MAF_OZ_SJ=4
run hints00 MAF_OZ_HINTS=00
run hints20 MAF_OZ_HINTS=20
run base_b MAF_OZ_GROUPS=2
