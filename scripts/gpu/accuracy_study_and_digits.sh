#!/bin/bash
# round 2, call A: full GPU test suite, smoke, accuracy study of the digit configurations, bench at three digit settings
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/a_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/a_pytest_gpu.log | head; grep -E "^E  " $O/a_pytest_gpu.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/a_smoke.log
rm -f $O/a_accuracy_study.jsonl
STUDY_OUT=$O/a_accuracy_study.jsonl timeout 1500 python scripts/accuracy_study.py > $O/a_accuracy_study.log 2>&1; echo "study rc=$?"; tail -2 $O/a_accuracy_study.log
for dg in 7,7 7,6 6,6; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --digits $dg > $O/a_bench_$dg.log 2>&1; echo "bench $dg rc=$?"
  tail -1 $O/a_bench_$dg.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print(d['roofline']['per_kernel_ms_per_step'])"
done
MAF_MU_ALPHA=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/a_bench_mualpha0.log 2>&1; echo "bench mu_alpha=0 rc=$?"
tail -1 $O/a_bench_mualpha0.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'])
print(d['roofline']['per_kernel_ms_per_step'])"
MAF_VBAR_TIGHT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/a_bench_tight1.log 2>&1; echo "bench tight rc=$?"
tail -1 $O/a_bench_tight1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'])
print(d['roofline']['per_kernel_ms_per_step'])"
