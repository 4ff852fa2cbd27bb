#!/bin/bash
TAG=${1:-fp16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gemm+netops"; timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_netops.py -m gpu -q -x -s > $OUT/pytest_gemm.log 2>&1; echo "rc=$?"; grep -E "relative L2|passed|failed|Error" $OUT/pytest_gemm.log | tail -8
echo "== parity full size"; timeout 600 python tools/gpu/parity_fullsize.py --no-oracle16 > $OUT/parity_fullsize.jsonl 2> $OUT/parity.err; echo "rc=$?"; cat $OUT/parity_fullsize.jsonl; tail -3 $OUT/parity.err
echo "== bench fp16"; AUDIOLAB_B200_NET_DTYPE=fp16 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > $OUT/bench_fp16.json 2> $OUT/bench_fp16.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_fp16.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])"; tail -3 $OUT/bench_fp16.err
echo "== bench bf16"; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_bf16.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])"
