#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x -rA --no-header -p no:cacheprovider > gpurun_out/pytest_ozaki.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|PASSED)|passed|failed" gpurun_out/pytest_ozaki.log | head -30
grep -E "^E  " gpurun_out/pytest_ozaki.log | head -12
timeout 300 python scripts/bench_gemm_i8.py 4096 32768 2>&1 | tee gpurun_out/gemm_i8.jsonl
