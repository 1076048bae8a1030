#!/bin/bash
# stored kernel derivatives (MAF_G_ROUTE): parity tests with the route on, bench A/B inside one call
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py -m gpu -q -x --no-header -p no:cacheprovider > $O/y_pytest_groute.log 2>&1; echo "pytest rc=$?"; tail -3 $O/y_pytest_groute.log; grep -E "^E  " $O/y_pytest_groute.log | head -5
for g in 1 0 1 0; do
  MAF_G_ROUTE=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/y_bench_g$g.log 2>&1; echo "bench g_route=$g rc=$?"
  tail -1 $O/y_bench_g$g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
