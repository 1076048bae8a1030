#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 5 6 7; do
  MAF_OZ_VARIANT=3 MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 32768 2>&1 | sed "s/^/v3 dbg$d /" | cut -c1-120 | tee -a gpurun_out/gemm_i8_v3dbg.jsonl
done
