#!/bin/bash
# A/B of the centred cross-covariance planes (MAF_CENTRE_K10): tests, switch study, bench, both settings
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for c in 1 0; do
  export MAF_CENTRE_K10=$c
  timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $O/u_pytest_gpu_c$c.log 2>&1; echo "centre=$c pytest rc=$?"; tail -2 $O/u_pytest_gpu_c$c.log; grep -E "^E  " $O/u_pytest_gpu_c$c.log | head -5
  rm -f $O/u_level_switch_study_c$c.jsonl
  STUDY_KERNELS=matern52 STUDY_OUT=$O/u_level_switch_study_c$c.jsonl timeout 600 python scripts/level_switch_study.py > $O/u_level_switch_study_c$c.log 2>&1; echo "study rc=$?"; grep "tol 2.5e-10" $O/u_level_switch_study_c$c.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/u_bench_c$c.log 2>&1; echo "bench rc=$?"
  tail -1 $O/u_bench_c$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d.get('level_switch') and (d['level_switch']['products_forward_avg'], d['level_switch']['products_reverse_avg'], d['level_switch']['calibrated_err_cov_abs']))
print(d['roofline']['per_kernel_ms_per_step'])"
done
