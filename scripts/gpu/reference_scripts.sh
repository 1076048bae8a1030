#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python scripts/run_reference_scripts_gpu.py > $O/n_reference_scripts.jsonl 2> $O/n_reference_scripts.err; echo "reference scripts rc=$?"; cat $O/n_reference_scripts.jsonl | cut -c1-400; tail -5 $O/n_reference_scripts.err
