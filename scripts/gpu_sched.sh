#!/bin/bash
# A/B of environment knobs on the bench workload: one line of env assignments per run
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sched.log 2>&1; echo "$* rc=$?"
  tail -1 gpurun_out/bench_sched.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['per_kernel_ms_per_step']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], ' '.join('%s %.2f' % (a, k[a]) for a in k))"
}
while read -r line; do [ -n "$line" ] && run $line; done <<CFG
MAF_CHUNK_ROWS=65536
MAF_CHUNK_ROWS=32768
MAF_CHUNK_ROWS=16384
MAF_CHUNK_ROWS=49152
MAF_CHUNK_ROWS=65536
MAF_CHUNK_ROWS=32768
CFG
