#!/bin/bash
# ncu --set full of the covariance kernels and the Vbar / posterior kernels (one chunk each), after a plain run
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --restarts 16384 --no-cpu-baseline"
$SMALL > gpurun_out/bench_small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'cross_fwd_planes|cross_bwd|vbar_planes|posterior_reg' -s 16 -c 4 -f -o gpurun_out/prof_cov_v9 $SMALL > gpurun_out/ncu_cov.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_cov.log
