#!/bin/bash
# full gpu suite (both contraction engines), bench (default = INT8), ncu launch list + full capture of the INT8 GEMM
mkdir -p gpurun_out
echo skip-pytest
tail -3 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head; grep -E "^E  " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 --gemm f64 --no-cpu-baseline > gpurun_out/bench_f64.log 2>&1; echo "bench f64 rc=$?"
tail -1 gpurun_out/bench_f64.log | cut -c1-300
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --steps 1 --warmup 3 --restarts 16384 --no-cpu-baseline"
$SMALL > gpurun_out/bench_small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_gemm_v4 -s 8 -c 2 -f -o gpurun_out/prof_i8v4 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
