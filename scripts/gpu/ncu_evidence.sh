#!/bin/bash
# ncu evidence: launch list of the bench command, one --set full capture of the contraction kernel and of the covariance
# kernels; the captures are summarised ON the box (raw page of the counters the design argues from) and removed:
# gpurun_out/ carries at most 64 MiB back.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-agent"
$BENCH > $O/p_bench_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/p_launches.csv $BENCH > $O/p_ncu_launches.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_ncu.py launches $O/p_launches.csv > $O/p_launches_summary.txt 2>&1
SMALL="python bench.py --steps 1 --warmup 3 --restarts 16384 --no-cpu-baseline --no-agent"
$SMALL > $O/p_bench_small_plain.log 2>&1; echo "small plain rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_gemm_v4 -s 24 -c 8 -f -o /tmp/p_prof_gemm $SMALL > $O/p_ncu_full_gemm.log 2>&1; echo "full capture gemm rc=$?"
python scripts/summarize_ncu.py full /tmp/p_prof_gemm.ncu-rep > $O/p_gemm_ncu_full_summary.txt 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cross_fwd_planes|cross_bwd|vbar_planes|posterior_reg" -s 24 -c 12 -f -o /tmp/p_prof_cov $SMALL > $O/p_ncu_full_cov.log 2>&1; echo "full capture cov rc=$?"
python scripts/summarize_ncu.py full /tmp/p_prof_cov.ncu-rep > $O/p_cov_ncu_full_summary.txt 2>&1
ls -la $O/p_*; head -12 $O/p_launches_summary.txt
