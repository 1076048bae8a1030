#!/bin/bash
# GPU session 2: parity tests, then ncu launch list of the bench command and one full capture of the
# dominant kernel.  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --steps 1 --warmup 3 --restarts 16384 --no-cpu-baseline"
$SMALL > gpurun_out/bench_small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_f64 -s 2 -c 2 -f -o gpurun_out/prof_gemm $SMALL > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-600
ls -la gpurun_out
