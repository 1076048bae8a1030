#!/bin/bash
mkdir -p gpurun_out
for sj in 4 2 5; do for h in 00 02 12 10; do
  MAF_OZ_HINTS=$h MAF_OZ_SJ=$sj timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | grep lower | sed "s/^/sj=$sj hints=$h /" | cut -c1-125 | tee -a gpurun_out/gemm_i8_hints2.jsonl
done; done
