#!/bin/bash
# One gpurun call: GPU tests, kernel roofline, bench line, smoke, bounded ncu launch list, ncu --set full of K1/K2.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_round2.sh [tag]'
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 420 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|error|relative error|SI-SDR" $OUT/pytest_gpu.log | tail -20
echo "== kernel_bench" ; timeout 150 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== bench" ; timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== smoke" ; timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/smoke.log
echo "== ncu launch list (bench step, 16 s track = one batch of chunks)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode --track-seconds 16 > $OUT/launches_bench.log 2>&1 ; echo "rc=$?"
echo "== ncu full (K1/K2 packed, kernel_bench --once)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'stft_pk2_kernel|istft_pk2_kernel' \
    -o $OUT/prof_pk -f python tools/kernel_bench.py --once --cases roformer_2048_441 > $OUT/prof_pk.log 2>&1 ; echo "rc=$?"
ncu -i $OUT/prof_pk.ncu-rep --page raw --csv > $OUT/prof_pk_raw.csv 2>/dev/null
ncu -i $OUT/prof_pk.ncu-rep --page details > $OUT/prof_pk_details.txt 2>/dev/null
ncu -i $OUT/prof_pk.ncu-rep --page source --csv -k regex:istft_pk2 > $OUT/prof_pk_istft_source.csv 2>/dev/null
ls -la $OUT
