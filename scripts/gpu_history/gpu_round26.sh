#!/bin/bash
mkdir -p gpurun_out
timeout 240 python scripts/check_i8_variants.py 4 > gpurun_out/check_v4.log 2>&1; echo "check rc=$?"
tail -25 gpurun_out/check_v4.log
for v in 3 4; do
  MAF_OZ_VARIANT=$v timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | sed "s/^/v$v /" | cut -c1-200 | tee -a gpurun_out/gemm_i8_v4.jsonl
done
