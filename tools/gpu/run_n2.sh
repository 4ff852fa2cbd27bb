#!/bin/bash
# N-GPU bench (default mode at N > 1 = chunk-range + halo exchange) with reduced cfg4 / cfg5 sizes.  Usage: run_n2.sh <tag> <N>
TAG=${1:-n2}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench --gpus $N (chunk-range default)"; AUDIOLAB_CFG4_SONGS_PER_GPU=4 AUDIOLAB_CFG5_MINUTES=12 timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; python - <<PY
import json
d=json.load(open('$OUT/bench_n$N.json'))
for k in ('value','ms_per_step','scaling','e2e','sharding','tracks_mode','gpu_launches'):
    print(k, json.dumps(d.get(k))[:700])
print(json.dumps(d.get('configs'))[:1500])
PY
tail -4 $OUT/bench_n$N.err
echo "== reference arm"; AUDIOLAB_REF_BUDGET_S=30 timeout 300 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ref_n$N.json
