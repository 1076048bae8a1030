#!/bin/bash
mkdir -p gpurun_out
for sj in 1 2 4 8 16 64; do
  MAF_OZ_SJ=$sj timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | grep lower | sed "s/^/sj=$sj /" | cut -c1-220 | tee -a gpurun_out/gemm_i8_sj.jsonl
done
