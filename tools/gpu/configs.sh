#!/bin/bash
# Last call of the round: GPU tests on the final tree + realtime factor of the other BASELINE configurations.
TAG=${1:-cfg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 120 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== config_bench" ; timeout 150 python tools/config_bench.py --budget-s 20 > $OUT/config_bench.jsonl 2> $OUT/config_bench.err ; echo "rc=$?" ; cat $OUT/config_bench.jsonl ; tail -5 $OUT/config_bench.err
