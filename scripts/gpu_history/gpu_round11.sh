#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_ozaki.log 2>&1; echo "pytest ozaki v3 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_ozaki.log | head
grep -E "^E  " gpurun_out/pytest_ozaki.log | head -12
for d in 0 4 6 5; do
  MAF_OZ_VARIANT=3 MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 32768 2>&1 | sed "s/^/v3 dbg$d /" | cut -c1-200 | tee -a gpurun_out/gemm_i8_v3.jsonl
done
