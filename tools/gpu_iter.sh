#!/bin/bash
# Intermediate gpurun call: GPU tests, kernel roofline, bench line, ncu --set full of the spectral / netops kernels.
# Usage: gpurun --timeout 600 -- 'bash tools/gpu_iter.sh [tag]'
TAG=${1:-it}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -5 $OUT/pytest_gpu.log
echo "== kernel_bench" ; timeout 150 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== kernel_bench K2 with 8 warps per CTA" ; AL_IP_WARPS=8 timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee $OUT/kernel_bench_ipw8.jsonl
echo "== bench" ; timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== ncu full: spectral kernels, kernel_bench --once"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'istft_pk2_kernel|ola_gather_kernel|resample_rb_kernel|istft_kernel|gelu_bf16' \
    -o $OUT/prof_k -f python tools/kernel_bench.py --once --cases roformer_2048_441,htdemucs_4096_1024 --gelu > $OUT/prof_k.log 2>&1 ; echo "rc=$?"
ncu -i $OUT/prof_k.ncu-rep --page raw --csv > $OUT/prof_k_raw.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:istft_pk2 > $OUT/prof_k_istft_source.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:'istft_kernel' > $OUT/prof_k_istft4_source.csv 2>/dev/null
rm -f $OUT/prof_k.ncu-rep
ls -la $OUT
