#!/bin/bash
# 2 GPUs: sharded suggest_inputs on NCCL against one process; weak / strong bench lines at N=2
mkdir -p gpurun_out/r2
O=gpurun_out/r2
SHARD_OUT=$O/l_sharded_suggest timeout 300 python scripts/check_sharded_suggest.py > $O/l_shard_1.log 2>&1; echo "1 process rc=$?"; tail -2 $O/l_shard_1.log
SHARD_OUT=$O/l_sharded_suggest timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/check_sharded_suggest.py > $O/l_shard_2.log 2>&1; echo "2 ranks rc=$?"; tail -2 $O/l_shard_2.log
python - <<'PY'
import json
a=json.load(open("gpurun_out/r2/l_sharded_suggest.1.json")); b=json.load(open("gpurun_out/r2/l_sharded_suggest.2.json"))
ok=all(x["index"]==y["index"] and x["x"]==y["x"] and x["loss"]==y["loss"] for x,y in zip(a,b))
print("1 process vs 2 ranks identical (suggestion, loss, index):", ok)
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > $O/l_bench_n2_weak.log 2>&1; echo "bench weak N=2 rc=$?"; tail -1 $O/l_bench_n2_weak.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > $O/l_bench_n2_strong.log 2>&1; echo "bench strong N=2 rc=$?"; tail -1 $O/l_bench_n2_strong.log | cut -c1-400
