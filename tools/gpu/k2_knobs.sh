#!/bin/bash
# pk5 token knobs: suspend-time hint of the token waits, relaxed token arrivals
TAG=${1:-k2knobs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "AL_IP_RING=2" "AL_IP_RING=3" "AL_IP_RING=3 AL_IP_HINT=100" "AL_IP_RING=3 AL_IP_HINT=1000" "AL_IP_RING=3 AL_IP_RELAXED=1" "AL_IP_RING=3 AL_IP_HINT=100 AL_IP_RELAXED=1"; do
  echo "== kernel_bench $v"; env $v timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee -a $OUT/kernel_bench_knobs.jsonl
done
echo "== pytest (AL_IP_RING=3 AL_IP_HINT=100 AL_IP_RELAXED=1)"; AL_IP_RING=3 AL_IP_HINT=100 AL_IP_RELAXED=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k istft > $OUT/pytest_knobs.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_knobs.log
