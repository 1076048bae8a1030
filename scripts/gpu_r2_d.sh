#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/e_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/e_pytest_gpu.log | head; grep -E "^E  " $O/e_pytest_gpu.log | head -20
run_bench () {
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/e_bench_$1.log 2>&1; echo "bench $1 rc=$?"
  tail -1 $O/e_bench_$1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d.get('level_switch') and (d['level_switch']['products_forward_avg'], d['level_switch']['products_reverse_avg']))
print(d['roofline']['per_kernel_ms_per_step'])"
}
run_bench lean
