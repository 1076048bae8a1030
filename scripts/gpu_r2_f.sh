#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/f_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/f_pytest_gpu.log | head; grep -E "^E  " $O/f_pytest_gpu.log | head -20
show () { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"; }
(cd _r1 && timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ../$O/f_bench_r1tree.log 2>&1; echo "bench r1 tree rc=$?"); show $O/f_bench_r1tree.log
MAF_LEVEL_SWITCH=0 MAF_MU_ALPHA=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/f_bench_plain.log 2>&1; echo "bench now, plain rc=$?"; show $O/f_bench_plain.log
MAF_LEVEL_SWITCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/f_bench_mualpha.log 2>&1; echo "bench now, mu_alpha rc=$?"; show $O/f_bench_mualpha.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/f_bench_default.log 2>&1; echo "bench now, default rc=$?"; show $O/f_bench_default.log
