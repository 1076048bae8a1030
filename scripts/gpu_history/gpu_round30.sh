#!/bin/bash
# A/B: register-stash epilogue (variant 5) and 6-plane reverse contraction
mkdir -p gpurun_out
MAF_OZ_VARIANT=5 timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -5
for cfg in "4 7" "5 7" "5 6" "4 6"; do
set -- $cfg
MAF_OZ_VARIANT=$1 MAF_OZ_NS_BWD=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$1_b$2.log 2>&1; echo "variant=$1 ns_bwd=$2 rc=$?"
tail -1 gpurun_out/bench_v$1_b$2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
