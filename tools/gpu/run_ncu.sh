#!/bin/bash
# ncu --set full captures exported to CSV ON THE BOX (the .ncu-rep files are deleted: gpurun_out/ is capped at 64 MiB).
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu gemm shapes"; timeout 420 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -c 8 -f -o /tmp/ncu_gemm python tools/gpu/gemm_ncu_target.py > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"; tail -1 $OUT/ncu_gemm.log
ncu -i /tmp/ncu_gemm.ncu-rep --page raw --csv > $OUT/ncu_gemm_raw.csv 2>/dev/null
ncu -i /tmp/ncu_gemm.ncu-rep --page source --csv > $OUT/ncu_gemm_source.csv 2>/dev/null
echo "== ncu generic kernels"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"istft_kernel|stft_kernel" -f -o /tmp/ncu_generic python tools/kernel_bench.py --once --cases htdemucs_4096_1024,mdx_6144_1024 > $OUT/ncu_generic.log 2>&1; echo "rc=$?"; tail -1 $OUT/ncu_generic.log
ncu -i /tmp/ncu_generic.ncu-rep --page raw --csv > $OUT/ncu_generic_raw.csv 2>/dev/null
ncu -i /tmp/ncu_generic.ncu-rep --page source --csv -k regex:'^(void )?(al::)?istft_kernel' > $OUT/ncu_generic_istft_source.csv 2>/dev/null
ncu -i /tmp/ncu_generic.ncu-rep --page source --csv -k regex:'^(void )?(al::)?stft_kernel' > $OUT/ncu_generic_stft_source.csv 2>/dev/null
for f in $OUT/ncu_generic_*_source.csv; do for i in 0 1 2 3; do python tools/ncu_phase_split.py $f $i 2>/dev/null; done > ${f%_source.csv}_phases.txt; done
gzip -f $OUT/*_source.csv
ls -la $OUT; du -sh gpurun_out
