#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python scripts/float32_mode_study.py > $O/q_float32_mode_study.jsonl 2> $O/q_float32_mode_study.err; echo "study rc=$?"; cat $O/q_float32_mode_study.jsonl | cut -c1-330; tail -3 $O/q_float32_mode_study.err
for dg in 5,5 4,4 3,3; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent --dtype float32 --digits $dg > $O/q_bench_f32_$dg.log 2>&1; echo "bench f32 $dg rc=$?"
  tail -1 $O/q_bench_f32_$dg.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
