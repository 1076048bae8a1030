#!/bin/bash
mkdir -p gpurun_out
for d in 0 6 4 2; do
  MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | sed "s/^/hint dbg$d /" | cut -c1-130 | tee -a gpurun_out/gemm_i8_hint.jsonl
done
