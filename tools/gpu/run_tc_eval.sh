#!/bin/bash
# GEMM timings + GPU tests + bench (tc path, with / without the band-axis attention kernel) [+ full-size parity table]
TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== gemm_debug"; timeout 300 python tools/gpu/gemm_debug.py > $OUT/gemm_debug.log 2>&1; echo "rc=$?"; grep kind $OUT/gemm_debug.log; grep -v kind $OUT/gemm_debug.log | grep -v '"bad": 0' | tail -5
echo "== pytest gemm+netops"; AUDIOLAB_B200_BAND_ATTN=1 timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_netops.py -m gpu -q > $OUT/pytest_gemm.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gemm.log
echo "== bench tc"; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'])"; tail -3 $OUT/bench.err
echo "== bench tc + band attention"; AUDIOLAB_B200_BAND_ATTN=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_band.json 2> $OUT/bench_band.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_band.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'])"; tail -3 $OUT/bench_band.err
if [ "$2" == "parity" ]; then
  echo "== parity full size"; timeout 600 python tools/gpu/parity_fullsize.py > $OUT/parity_fullsize.jsonl 2> $OUT/parity.err; echo "rc=$?"; cat $OUT/parity_fullsize.jsonl; tail -3 $OUT/parity.err
fi
