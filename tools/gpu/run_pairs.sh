#!/bin/bash
# A/B of CTA pairs (tcgen05 cta_group::2) for the 256-wide EPI_BF16 / EPI_GLU GEMM tiles
TAG=${1:-pairs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for pv in 1 0; do
  echo "== gemm_debug PAIRS=$pv"; AL_GEMM_PAIRS=$pv timeout 300 python tools/gpu/gemm_debug.py > $OUT/gemm_debug_pairs$pv.log 2>&1; echo "rc=$?"
  grep kind $OUT/gemm_debug_pairs$pv.log | grep -v res; grep -v kind $OUT/gemm_debug_pairs$pv.log | grep -v '"bad": 0' | cut -c1-600 | tail -5
done
echo "== pytest gemm+netops"; timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_netops.py -m gpu -q -x > $OUT/pytest_gemm.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gemm.log
for pv in 1 0; do
  echo "== bench PAIRS=$pv"; AL_GEMM_PAIRS=$pv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > $OUT/bench_pairs$pv.json 2> $OUT/bench_pairs$pv.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('$OUT/bench_pairs$pv.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'), {k:(round(v['ms'],1),round(v['tflops'],0)) for k,v in d['kernels']['al_gemm_bf16'].items()})
PY
  tail -3 $OUT/bench_pairs$pv.err
done
