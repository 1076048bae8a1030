#!/bin/bash
# ncu --set full of the non-GEMM kernels of one chunk (plain run first)
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --restarts 8192 --no-cpu-baseline"
$SMALL > gpurun_out/bench_small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cross_bwd|cross_fwd_planes|vbar_planes|posterior_reg|mc_acq" -s 20 -c 5 -f -o gpurun_out/prof_small $SMALL > gpurun_out/ncu_small.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_small.log
