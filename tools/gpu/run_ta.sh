#!/bin/bash
# in-situ A/B of the time-axis attention kernel (AUDIOLAB_B200_TIME_ATTN=1) against cuDNN SDPA + gate pass
TAG=${1:-ta}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python tools/gpu/fattn_debug.py 2>&1 | grep kind
for ta in 1 0; do
  echo "== bench TIME_ATTN=$ta"; AUDIOLAB_B200_TIME_ATTN=$ta timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > $OUT/bench_ta$ta.json 2> $OUT/bench_ta$ta.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('$OUT/bench_ta$ta.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'].get('frac'), d['clocks'])
PY
  tail -3 $OUT/bench_ta$ta.err
done
AUDIOLAB_B200_TIME_ATTN=1 timeout 600 python tools/gpu/parity_fullsize.py 2>&1 | tail -4
