#!/bin/bash
# Round-end evidence call (1 GPU): GPU tests, kernel roofline, default bench line, smoke, ncu --set full of al_istft inside
# a timed bench step (traffic), ncu launch list of one timed bench step of the default configuration.
# Usage: gpurun --timeout 600 -- 'bash tools/gpu/evidence.sh [tag]'
TAG=${1:-fin}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== kernel_bench" ; timeout 120 python tools/kernel_bench.py --gelu > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== bench" ; timeout 200 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== smoke" ; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/smoke.log
echo "== ncu full: al_istft inside the bench step (traffic)"
timeout 120 ncu --set full --clock-control none --profile-from-start off -k regex:'istft_pk[234]_kernel' -c 1 -o $OUT/prof_bench_istft -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode --track-seconds 60 > $OUT/prof_bench_istft.log 2>&1 ; echo "rc=$?"
ncu -i $OUT/prof_bench_istft.ncu-rep --page raw --csv > $OUT/prof_bench_istft_raw.csv 2>/dev/null
rm -f $OUT/prof_bench_istft.ncu-rep
echo "== ncu launch list (one timed bench step, default configuration: 60 s track, 27 chunks per call)"
timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode --track-seconds 60 > $OUT/launches_bench.log 2>&1 ; echo "rc=$?"
ls -la $OUT
