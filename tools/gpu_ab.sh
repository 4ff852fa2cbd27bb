#!/bin/bash
# Short A/B call: GPU tests, kernel roofline with the default variants, K2 preload depth and OLA rows-per-thread A/B, bench.
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 100 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== kernel_bench (defaults: AL_IP_PRE=2, two rows per thread in al_ola_gather)" ; timeout 60 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
for pre in 0 3; do echo "== K2 AL_IP_PRE=$pre" ; AL_IP_PRE=$pre timeout 40 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee $OUT/kernel_bench_pre$pre.jsonl ; done
echo "== OLA one row per thread" ; AL_OLA_RB1=1 timeout 40 python tools/kernel_bench.py --only ola 2>&1 | tee $OUT/kernel_bench_ola_rb1.jsonl
echo "== bench" ; timeout 120 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
