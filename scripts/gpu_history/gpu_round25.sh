#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_ozaki.log 2>&1; echo "pytest ozaki rc=$?"
tail -2 gpurun_out/pytest_ozaki.log
for d in 0 6 4; do
  MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | sed "s/^/epi2 dbg$d /" | cut -c1-130 | tee -a gpurun_out/gemm_i8_epi2.jsonl
done
