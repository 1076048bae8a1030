#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_variants2.jsonl
for v in 0 5 6 7; do
  echo "variant $v" | tee -a gpurun_out/gemm_variants2.jsonl; MAF_GEMM_VARIANT=$v timeout 300 python scripts/bench_gemm.py 4096 32768 2>&1 | tee -a gpurun_out/gemm_variants2.jsonl
done
bash scripts/gpu_profile.sh
