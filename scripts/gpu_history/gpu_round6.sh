#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./scripts/microbench_tcgen05 > gpurun_out/microbench_tcgen05.jsonl 2>&1; echo "tcgen05 microbench rc=$?"
cat gpurun_out/microbench_tcgen05.jsonl
bash scripts/gpu_tests.sh
