#!/bin/bash
# ncu --set full capture of one kernel from tools/kernel_bench.py; exports small CSV/txt pages.
# usage: bash tools/gpu/prof_kernel.sh TAG KERNEL_REGEX CASE ONLY
TAG=$1; KRE=$2; CASE=${3:-roformer_2048_441}; ONLY=${4:-stft}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 1 -c 1 -o $OUT/prof -f \
   python tools/kernel_bench.py --once --cases $CASE --only $ONLY > $OUT/prof.log 2>&1 ; echo "ncu full rc=$?"
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page details > $OUT/prof_details.txt 2>/dev/null
ls -la $OUT/prof.ncu-rep
[ $(stat -c %s $OUT/prof.ncu-rep) -gt 30000000 ] && rm -f $OUT/prof.ncu-rep
du -sh $OUT
