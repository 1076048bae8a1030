#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_ozaki.log 2>&1; echo "pytest ozaki rc=$?"
tail -2 gpurun_out/pytest_ozaki.log
for pf in 1 0; do for d in 0 2; do
  MAF_OZ_PF=$pf MAF_OZ_DBG=$d timeout 120 python scripts/bench_gemm_i8.py 4096 65536 2>&1 | sed "s/^/pf=$pf dbg$d /" | cut -c1-130 | tee -a gpurun_out/gemm_i8_pf.jsonl
done; done
