#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do for d in 4 5 6 7; do
  MAF_OZ_VARIANT=$v MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 32768 2>&1 | sed "s/^/v$v dbg$d /" | cut -c1-130 | tee -a gpurun_out/gemm_i8_dbg.jsonl
done; done
