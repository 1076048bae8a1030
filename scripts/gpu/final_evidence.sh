#!/bin/bash
# final state of the round: full GPU suite, smoke, bench (with CPU baseline and agent timings), reference arm, BASELINE configs
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider > $O/z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/z_pytest_gpu.log; grep -E "^(FAILED|ERROR)" $O/z_pytest_gpu.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $O/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/z_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/z_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/z_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/z_bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -1 $O/z_bench_reference.log | cut -c1-200
timeout 600 python scripts/run_configs.py > $O/z_configs_C2_C5.jsonl 2> $O/z_configs.err; echo "configs rc=$?"; cut -c1-260 $O/z_configs_C2_C5.jsonl; tail -2 $O/z_configs.err
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-agent"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/z_launches.csv $BENCH > $O/z_ncu_launches.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_ncu.py launches $O/z_launches.csv > $O/z_launches_summary.txt 2>&1; head -14 $O/z_launches_summary.txt
