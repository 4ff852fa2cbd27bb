#!/bin/bash
# Final short call: GPU tests, kernel roofline (defaults), K2 interior overlap-add A/B, contract line.
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 100 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== kernel_bench (defaults)" ; timeout 60 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== K2 AL_IP_OLAFAST=0" ; AL_IP_OLAFAST=0 timeout 40 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee $OUT/kernel_bench_olafast0.jsonl
echo "== bench" ; timeout 120 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
