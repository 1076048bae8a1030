#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/c_pytest_gpu.log | head; grep -E "^E  " $O/c_pytest_gpu.log | head -20
rm -f $O/c_level_switch_study.jsonl
STUDY_OUT=$O/c_level_switch_study.jsonl timeout 1500 python scripts/level_switch_study.py > $O/c_level_switch_study.log 2>&1; echo "study rc=$?"; grep "switch 1 tol 2.5e-10" $O/c_level_switch_study.log
for sw in 1 0; do
  MAF_LEVEL_SWITCH=$sw timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/c_bench_sw$sw.log 2>&1; echo "bench switch=$sw rc=$?"
  tail -1 $O/c_bench_sw$sw.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['adam_loop']['value'], d['clocks'], d.get('level_switch'))
print(d['roofline']['per_kernel_ms_per_step'])"
done
