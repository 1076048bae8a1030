#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
C4_MODE=pending timeout 240 python scripts/run_c4.py > $O/k_c4_pending.jsonl 2> $O/k_c4_pending.err; echo "c4 pending rc=$?"; tail -1 $O/k_c4_pending.jsonl; tail -3 $O/k_c4_pending.err
MAF_PEND_DEDUPE=0 C4_MODE=pending timeout 240 python scripts/run_c4.py > $O/k_c4_pending_nodedupe.jsonl 2> $O/k_c4_pending_nodedupe.err; echo "c4 pending (no dedupe) rc=$?"; tail -1 $O/k_c4_pending_nodedupe.jsonl; tail -3 $O/k_c4_pending_nodedupe.err
C4_MODE=incremental timeout 240 python scripts/run_c4.py > $O/k_c4_incremental.jsonl 2> $O/k_c4_incremental.err; echo "c4 incremental rc=$?"; tail -1 $O/k_c4_incremental.jsonl; tail -3 $O/k_c4_incremental.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/k_bench_default.log 2>&1; echo "bench default rc=$?"
tail -1 $O/k_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['agent_suggest_inputs'])"
