#!/bin/bash
# A/B of K2 variants in one call.  VARIANTS = ';'-separated env settings; CAND = the one the GPU tests run with
TAG=${1:-k2ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CAND=${CAND:-AL_IP_RING=2 AL_IP_ZREG=1}
VARIANTS=${VARIANTS:-AL_IP_RING=1;AL_IP_RING=2;AL_IP_RING=2 AL_IP_ZREG=1;AL_IP_RING=3;AL_IP_RING=2;AL_IP_RING=2 AL_IP_ZREG=1}
echo "== pytest ($CAND)"; env $CAND timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_demix_gpu.py -m gpu -q -x > $OUT/pytest_cand.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_cand.log
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  echo "== kernel_bench $v"; env $v timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | sed "s/^{/{\"variant\": \"$v\", /" | tee -a $OUT/kernel_bench_ab.jsonl
done
