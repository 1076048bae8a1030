#!/bin/bash
# round-1 v9 evidence: GPU parity suite, smoke, bench both arms, launch list,
# ncu --set full of the INT8 contraction.  Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --steps 1 --warmup 3 --restarts 16384 --no-cpu-baseline"
$SMALL > gpurun_out/bench_small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_gemm_v4 -s 12 -c 2 -f -o gpurun_out/prof_gemm_v9 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
