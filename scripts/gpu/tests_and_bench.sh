#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/i_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/i_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/i_pytest_gpu.log | head; grep -E "^E  " $O/i_pytest_gpu.log | head -20
show () { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['e2e']['value'], d['adam_loop']['value'])
print(d['roofline']['per_kernel_ms_per_step'])"; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/i_bench_default.log 2>&1; echo "bench default rc=$?"; show $O/i_bench_default.log
