#!/bin/bash
# N GPUs ($1): strong-scaling (BASELINE config 5) and weak-scaling bench lines
N=$1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N"; fi
timeout 300 $L bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --no-agent > $O/m_bench_n${N}_strong.log 2>&1; echo "bench strong N=$N rc=$?"; tail -1 $O/m_bench_n${N}_strong.log | cut -c1-330
if [ "$2" = "weak" ]; then
timeout 300 $L bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-agent > $O/m_bench_n${N}_weak.log 2>&1; echo "bench weak N=$N rc=$?"; tail -1 $O/m_bench_n${N}_weak.log | cut -c1-330
fi
