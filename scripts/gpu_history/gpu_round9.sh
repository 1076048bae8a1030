#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_ozaki.log 2>&1; echo "pytest ozaki v2 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_ozaki.log | head
grep -E "^E  " gpurun_out/pytest_ozaki.log | head -12
for v in 1 2; do
  MAF_OZ_VARIANT=$v timeout 300 python scripts/bench_gemm_i8.py 4096 32768 2>&1 | sed "s/^/v$v /" | tee -a gpurun_out/gemm_i8_v.jsonl
done
