#!/bin/bash
# issuer / epilogue cycle counters of the CTA-pair kernel, old (4) vs register-stash (5) epilogue (MAF_OZ_DIAG build)
mkdir -p gpurun_out
for v in 4 5; do
  MAF_OZ_STATS=1 MAF_OZ_VARIANT=$v timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | sed "s/^/v$v /" | cut -c1-600 | tee -a gpurun_out/gemm_i8_v5_stats.txt
done
