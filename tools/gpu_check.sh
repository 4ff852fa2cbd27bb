#!/bin/bash
# One gpurun call without the heavy ncu --set full pass: GPU tests, kernel roofline, bench line, smoke, ncu launch list.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -5 $OUT/pytest_gpu.log
echo "== kernel_bench" ; timeout 300 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== bench" ; timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -3 $OUT/bench.err
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/smoke.log
echo "== ncu launch list (bench step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode > $OUT/launches_bench.log 2>&1 ; echo "rc=$?"
ls -la $OUT
