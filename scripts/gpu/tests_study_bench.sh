#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
T=${1:-t}
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest_gpu.log; grep -E "^E  " $O/${T}_pytest_gpu.log | head
rm -f $O/${T}_level_switch_study.jsonl
STUDY_KERNELS=matern52 STUDY_OUT=$O/${T}_level_switch_study.jsonl timeout 600 python scripts/level_switch_study.py > $O/${T}_level_switch_study.log 2>&1; echo "study rc=$?"; grep "tol 2.5e-10" $O/${T}_level_switch_study.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/${T}_bench_default.log 2>&1; echo "bench default rc=$?"
tail -1 $O/${T}_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d.get('level_switch') and (d['level_switch']['products_forward_avg'], d['level_switch']['products_reverse_avg'], d['level_switch']['calibrated_err_cov_abs']))
print(d['roofline']['per_kernel_ms_per_step'])"
