#!/bin/bash
# whole GPU suite (incl. the full-size parity test), kernel_bench after the generic-kernel changes, default-ish bench
TAG=${1:-ev4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "SI-SDR|relative L2|passed|failed" $OUT/pytest_gpu.log | tail -8
echo "== kernel_bench"; timeout 200 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err; echo "rc=$?"; cat $OUT/kernel_bench.jsonl; tail -3 $OUT/kernel_bench.err
echo "== bench (no configs)"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg1,cfg3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python - <<PY
import json
d=json.load(open('$OUT/bench.json'))
for k in ('value','ms_per_step','dtype','e2e','precision_modes','gpu_launches'):
    print(k, json.dumps(d.get(k))[:500])
print(json.dumps(d.get('configs'))[:1500])
PY
tail -3 $OUT/bench.err
