#!/bin/bash
# A/B of the K2 kernels in one call: GPU tests with the candidate kernel (AL_IP_RING=$CAND), isolated timing of the variants
TAG=${1:-k2ab}; CAND=${CAND:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (AL_IP_RING=$CAND)"; AL_IP_RING=$CAND timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_demix_gpu.py -m gpu -q -x > $OUT/pytest_ring$CAND.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_ring$CAND.log
for v in "AL_IP_RING=1" "AL_IP_RING=2" "AL_IP_RING=3" "AL_IP_RING=2" "AL_IP_RING=3"; do
  echo "== kernel_bench $v"; env $v timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee -a $OUT/kernel_bench_ab.jsonl
done
