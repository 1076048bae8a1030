#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider -x > $O/j_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/j_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/j_pytest_gpu.log | head; grep -E "^E  " $O/j_pytest_gpu.log | head -20
show () { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['e2e']['value'], d['adam_loop']['value'])
print(d['roofline']['per_kernel_ms_per_step'])"; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j_bench_default.log 2>&1; echo "bench default rc=$?"; show $O/j_bench_default.log
C4_MODE=pending timeout 900 python scripts/run_c4.py > $O/j_c4_pending.jsonl 2> $O/j_c4_pending.err; echo "c4 pending rc=$?"; tail -1 $O/j_c4_pending.jsonl; tail -3 $O/j_c4_pending.err
MAF_PEND_DEDUPE=0 C4_MODE=pending timeout 900 python scripts/run_c4.py > $O/j_c4_pending_nodedupe.jsonl 2> $O/j_c4_pending_nodedupe.err; echo "c4 pending (no dedupe) rc=$?"; tail -1 $O/j_c4_pending_nodedupe.jsonl; tail -3 $O/j_c4_pending_nodedupe.err
C4_MODE=incremental timeout 900 python scripts/run_c4.py > $O/j_c4_incremental.jsonl 2> $O/j_c4_incremental.err; echo "c4 incremental rc=$?"; tail -1 $O/j_c4_incremental.jsonl; tail -3 $O/j_c4_incremental.err
