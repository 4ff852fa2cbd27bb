#!/bin/bash
# ncu --set full of the K2 ring kernel (istft_pk3) out of tools/kernel_bench.py: raw metrics, sync-split of the SASS view,
# hot SASS lines; plus the isolated kernel roofline of the tree.
# usage: gpurun --timeout 600 -- 'bash tools/gpu/prof_k2.sh TAG'
TAG=${1:-k2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -n "$SKIPKB" ] || { echo "== kernel_bench"; timeout 150 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err; echo "rc=$?"; cat $OUT/kernel_bench.jsonl; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"${KRE:-istft_pk3_kernel}" -s 1 -c 1 -o /tmp/prof_k2 -f \
   python tools/kernel_bench.py --once --cases roformer_2048_441 --only istft > $OUT/prof.log 2>&1 ; echo "ncu full rc=$?"
ncu -i /tmp/prof_k2.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i /tmp/prof_k2.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
python tools/ncu_raw_extract.py $OUT/prof_raw.csv > $OUT/prof_raw_summary.txt 2>&1; cat $OUT/prof_raw_summary.txt
python tools/ncu_sync_split.py $OUT/prof_source.csv istft_pk 0 0.5 > $OUT/prof_sync_split.txt 2>&1; cat $OUT/prof_sync_split.txt
python tools/ncu_hot_sass.py $OUT/prof_source.csv istft_pk 0 40 2>/dev/null | head -40 > $OUT/prof_hot_sass.txt; cat $OUT/prof_hot_sass.txt
gzip -f $OUT/prof_source.csv
du -sh $OUT
