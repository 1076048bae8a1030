#!/bin/bash
# default-mode parity, INT8-mode parity over the whole gpu suite, bench in INT8 mode, one full ncu capture of the i8 GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest f64 rc=$?"
tail -3 gpurun_out/pytest_gpu.log
MAF_GEMM_I8=1 timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_i8.log 2>&1; echo "pytest i8 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_i8.log | head -30
timeout 600 python bench.py --steps 3 --warmup 3 --gemm i8 --no-cpu-baseline > gpurun_out/bench_i8.log 2>&1; echo "bench i8 rc=$?"
tail -1 gpurun_out/bench_i8.log | cut -c1-1500
G="python scripts/bench_gemm_i8.py 4096 4096"
$G > gpurun_out/gemm_i8_small.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_gemm -c 2 -f -o gpurun_out/prof_i8 $G > gpurun_out/ncu_i8.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_i8.log
