#!/bin/bash
# last call of a round: every GPU test, smoke, the default contract line
TAG=${1:-fin}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -6 $OUT/pytest_gpu.log
echo "== smoke" ; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/smoke.log
echo "== bench" ; timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cut -c1-300 $OUT/bench.json ; tail -3 $OUT/bench.err
