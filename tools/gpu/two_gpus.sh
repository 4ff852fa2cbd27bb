#!/bin/bash
# 2-GPU gpurun call: NCCL halo-exchange parity check and the two multi-GPU bench modes.
# Usage: gpurun --gpus 2 --timeout 420 -- 'bash tools/gpu/two_gpus.sh [tag]'
TAG=${1:-n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== pytest -m gpu (rank-less, GPU 0)" ; timeout 200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== kernel_bench stft" ; timeout 100 python tools/kernel_bench.py --only stft 2>&1 | tee $OUT/kernel_bench_stft.jsonl
echo "== nccl shard check" ; timeout 150 $TR --master-port 29511 tools/nccl_shard_check.py > $OUT/nccl_shard_check.json 2> $OUT/nccl_shard_check.err ; echo "rc=$?" ; cat $OUT/nccl_shard_check.json ; tail -3 $OUT/nccl_shard_check.err
echo "== bench --gpus 2 (track per GPU)" ; timeout 150 $TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/bench_n2_tracks.json 2> $OUT/bench_n2_tracks.err ; echo "rc=$?" ; cut -c1-330 $OUT/bench_n2_tracks.json ; tail -3 $OUT/bench_n2_tracks.err
echo "== bench --gpus 2 --mode chunk-range" ; timeout 150 $TR --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --mode chunk-range > $OUT/bench_n2_chunkrange.json 2> $OUT/bench_n2_chunkrange.err ; echo "rc=$?" ; cut -c1-330 $OUT/bench_n2_chunkrange.json ; tail -3 $OUT/bench_n2_chunkrange.err
