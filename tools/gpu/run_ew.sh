#!/bin/bash
# A/B of the GEMM epilogue width (AL_GEMM_EW=8: two warps per TMEM lane quadrant, default: four) + GPU tests + bench
TAG=${1:-ew}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for ew in 16 8; do
  echo "== gemm_debug EW=$ew"; AL_GEMM_EW=$ew timeout 300 python tools/gpu/gemm_debug.py > $OUT/gemm_debug_ew$ew.log 2>&1; echo "rc=$?"
  grep kind $OUT/gemm_debug_ew$ew.log; grep -v kind $OUT/gemm_debug_ew$ew.log | grep -v '"bad": 0' | tail -5
done
echo "== pytest gemm+netops"; timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_netops.py -m gpu -q > $OUT/pytest_gemm.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gemm.log
for ew in 16 8; do
  echo "== bench EW=$ew"; AL_GEMM_EW=$ew timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > $OUT/bench_ew$ew.json 2> $OUT/bench_ew$ew.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('$OUT/bench_ew$ew.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'].get('frac'), {k:(round(v['ms'],1),round(v['tflops'],0)) for k,v in d['kernels']['al_gemm_bf16'].items()})
PY
  tail -3 $OUT/bench_ew$ew.err
done
