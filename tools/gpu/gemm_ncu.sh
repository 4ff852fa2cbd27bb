#!/bin/bash
# ncu --set full + source page of the to_qkv and FF Linear-1 launches of al_gemm_bf16 (tools/gpu/gemm_ncu_target.py)
TAG=${1:-gemm_ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 420 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -c 4 -f -o /tmp/ncu_gemm python tools/gpu/gemm_ncu_target.py > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"
ncu -i /tmp/ncu_gemm.ncu-rep --page raw --csv > $OUT/ncu_gemm_raw.csv 2>/dev/null
ncu -i /tmp/ncu_gemm.ncu-rep --page source --csv > $OUT/ncu_gemm_source.csv 2>/dev/null
python tools/ncu_raw_extract.py $OUT/ncu_gemm_raw.csv | grep -E "^==|time_duration|issue_active|stall samples|dram__bytes|registers" 
for i in 0 2; do python tools/ncu_sync_split.py $OUT/ncu_gemm_source.csv gemm_bf16 $i 0.8; python tools/ncu_hot_sass.py $OUT/ncu_gemm_source.csv gemm_bf16 $i 14; done
gzip -f $OUT/ncu_gemm_source.csv
