#!/bin/bash
mkdir -p gpurun_out
for h in 00 21 20 02 12 01 22; do for sj in 8 16; do
  MAF_OZ_HINTS=$h MAF_OZ_SJ=$sj timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | grep lower | sed "s/^/hints=$h sj=$sj /" | cut -c1-130 | tee -a gpurun_out/gemm_i8_hints.jsonl
done; done
