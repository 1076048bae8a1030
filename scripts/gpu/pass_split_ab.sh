#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
run_bench () {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/r_bench_$1.log 2>&1; echo "bench $1 rc=$?"
  tail -1 $O/r_bench_$1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
print(d['roofline']['per_kernel_ms_per_step'])"
}
run_bench split_2_4
MAF_NVCC_DEFS="-DOZ3_N0_6=3" python -m maximizingacquisitionfunctions_b200.build --force > $O/r_rebuild.log 2>&1; echo "rebuild rc=$?"
run_bench split_3_3
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -x -k "ozaki or bench_workload or level_switch or reference_code_golden" > $O/r_pytest_split33.log 2>&1; echo "pytest (3|3 build) rc=$?"; tail -2 $O/r_pytest_split33.log
