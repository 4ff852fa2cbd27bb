#!/bin/bash
# ncu --set full of the K1 packed kernel (stft_pk2) out of tools/kernel_bench.py: raw metrics, sync-split, hot SASS
TAG=${1:-k1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stft_pk2_kernel" -s 1 -c 1 -o /tmp/prof_k1 -f \
   python tools/kernel_bench.py --once --cases roformer_2048_441 --only stft > $OUT/prof.log 2>&1 ; echo "ncu full rc=$?"
ncu -i /tmp/prof_k1.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
python tools/ncu_raw_extract.py $OUT/prof_raw.csv > $OUT/prof_raw_summary.txt 2>&1; cat $OUT/prof_raw_summary.txt
python tools/ncu_sync_split.py $OUT/prof_source.csv stft_pk2 0 0.5 > $OUT/prof_sync_split.txt 2>&1; cat $OUT/prof_sync_split.txt
python tools/ncu_hot_sass.py $OUT/prof_source.csv stft_pk2 0 30 2>/dev/null > $OUT/prof_hot_sass.txt; cat $OUT/prof_hot_sass.txt
gzip -f $OUT/prof_source.csv
