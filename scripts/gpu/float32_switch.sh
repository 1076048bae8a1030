#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $O/s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/s_pytest_gpu.log; grep -E "^E  " $O/s_pytest_gpu.log | head
for sw in 1 0; do
  MAF_LEVEL_SWITCH=$sw timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent --dtype float32 > $O/s_bench_f32_sw$sw.log 2>&1; echo "bench f32 switch=$sw rc=$?"
  tail -1 $O/s_bench_f32_sw$sw.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d.get('level_switch') and (d['level_switch']['products_forward_avg'], d['level_switch']['products_reverse_avg']))
print(d['roofline']['per_kernel_ms_per_step'])"
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/s_bench_default.log 2>&1; echo "bench default rc=$?"
tail -1 $O/s_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['agent_suggest_inputs'])"
