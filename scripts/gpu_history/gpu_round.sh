#!/bin/bash
# One GPU session: peaks, parity tests, kernel bench, headline bench, smoke.  Everything is logged
# under gpurun_out/ and every stage runs under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.csv 2>&1
timeout 300 ./scripts/microbench_fp64 > gpurun_out/microbench.jsonl 2>&1; echo "microbench rc=$?"
timeout 900 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/gemm.jsonl 2>&1; echo "gemm bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
cat gpurun_out/microbench.jsonl gpurun_out/gemm.jsonl; tail -3 gpurun_out/bench.log; cat gpurun_out/smoke.log | tail -3
