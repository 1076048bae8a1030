#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"
tail -1 gpurun_out/bench_n$N.log | cut -c1-500
grep -i "error\|Traceback" gpurun_out/bench_n$N.log | head -5
