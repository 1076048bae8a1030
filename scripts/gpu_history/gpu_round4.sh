#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_variants.jsonl
for v in 0 4 5 6 7; do
  echo "variant $v" | tee -a gpurun_out/gemm_variants.jsonl; MAF_GEMM_VARIANT=$v timeout 300 python scripts/bench_gemm.py 4096 32768 2>&1 | tee -a gpurun_out/gemm_variants.jsonl
done
for v in 4 5; do MAF_GEMM_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or full_size or value_and_grad" -p no:cacheprovider 2>&1 | tail -2; done
timeout 600 python -m pytest tests/test_gpu_mirror.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
