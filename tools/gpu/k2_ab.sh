#!/bin/bash
# A/B of the K2 kernels in one call: GPU tests with the token-ordered kernel, isolated timing of pk3 / pk4 (11 and 7 consumers)
TAG=${1:-k2ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (AL_IP_RING=2)"; AL_IP_RING=2 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_demix_gpu.py -m gpu -q -x > $OUT/pytest_ring2.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_ring2.log

for v in "AL_IP_RING=1" "AL_IP_RING=2" "AL_IP_RING=1" "AL_IP_RING=2"; do
  echo "== kernel_bench $v"; env $v timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee -a $OUT/kernel_bench_ab.jsonl
done
