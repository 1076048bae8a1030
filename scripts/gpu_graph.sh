#!/bin/bash
# replayed optimizer steps (CUDA graph): parity tests, then loop latency on the small configurations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py -m gpu -q -x --no-header -p no:cacheprovider -k "adam or gd_and or replayed or sgd or c1 or c2 or halving" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_graph.log
timeout 600 python scripts/bench_adam_small.py > gpurun_out/adam_small.jsonl 2> gpurun_out/adam_small.err; echo "bench rc=$?"
cat gpurun_out/adam_small.jsonl; tail -5 gpurun_out/adam_small.err
