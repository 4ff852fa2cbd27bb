#!/bin/bash
# 2-GPU gpurun call on the current tree: NCCL halo-exchange parity check, the contract launch at N = 2 (default mode = chunk-range
# sharding + halo exchange, with its bitwise self-check) and the reference arm at N = 2.
# Usage: gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu/n2.sh [tag]'
TAG=${1:-n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== nccl shard check" ; timeout 150 $TR --master-port 29511 tools/nccl_shard_check.py > $OUT/nccl_shard_check.json 2> $OUT/nccl_shard_check.err ; echo "rc=$?" ; cat $OUT/nccl_shard_check.json ; tail -3 $OUT/nccl_shard_check.err
echo "== bench --gpus 2 (driver launch line)" ; timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err ; echo "rc=$?" ; cut -c1-400 $OUT/bench_n2.json ; tail -3 $OUT/bench_n2.err
