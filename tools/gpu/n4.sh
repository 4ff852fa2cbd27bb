#!/bin/bash
# N-GPU gpurun call (N = 4 or 8): the contract launch through the driver's launch line (chunk-range sharding + halo exchange)
N=${2:-4}; TAG=${1:-n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench --gpus $N" ; timeout 500 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ; echo "rc=$?" ; cut -c1-300 $OUT/bench_n$N.json ; tail -3 $OUT/bench_n$N.err
