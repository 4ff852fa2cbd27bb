#!/bin/bash
# ncu --set full of the four mask-network GEMM shapes; full-size parity table; isolated kernel roofline; ncu (with source)
# of the generic n_fft 4096 / 6144 kernels.
TAG=${1:-ev3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu gemm shapes"; timeout 420 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -c 8 -f -o $OUT/ncu_gemm python tools/gpu/gemm_ncu_target.py > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_gemm.log
echo "== parity full size"; timeout 600 python tools/gpu/parity_fullsize.py > $OUT/parity_fullsize.jsonl 2> $OUT/parity.err; echo "rc=$?"; cat $OUT/parity_fullsize.jsonl; tail -3 $OUT/parity.err
echo "== kernel_bench"; timeout 200 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err; echo "rc=$?"; cat $OUT/kernel_bench.jsonl; tail -3 $OUT/kernel_bench.err
echo "== ncu generic kernels"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"istft_kernel|stft_kernel" -f -o $OUT/ncu_generic python tools/kernel_bench.py --once --cases htdemucs_4096_1024,mdx_6144_1024 > $OUT/ncu_generic.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_generic.log
ls -la $OUT
