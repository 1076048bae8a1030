#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_parity.py -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
for v in 0 1 2 3; do
  echo "variant $v"; MAF_GEMM_VARIANT=$v timeout 300 python scripts/bench_gemm.py 4096 32768 2>&1 | tee -a gpurun_out/gemm_variants.jsonl
done
