#!/bin/bash
# One gpurun call: GPU tests, kernel roofline (+ K2 A/B of warps per CTA), bench line, smoke, ncu launch list of one
# timed bench step, ncu --set full of the spectral kernels (one launch each) and of al_istft inside the bench step.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu/full_profile.sh [tag]'
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -5 $OUT/pytest_gpu.log
echo "== kernel_bench" ; timeout 150 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
echo "== kernel_bench K2 with 8 warps per CTA" ; AL_IP_WARPS=8 timeout 100 python tools/kernel_bench.py --only istft --cases roformer_2048_441 2>&1 | tee $OUT/kernel_bench_ipw8.jsonl
echo "== bench, 27 chunks per mask-net call" ; timeout 200 python bench.py --steps 2 --warmup 3 --batch 27 --no-cpu-baseline > $OUT/bench_b27.json 2> $OUT/bench_b27.err ; echo "rc=$?" ; cut -c1-400 $OUT/bench_b27.json ; tail -3 $OUT/bench_b27.err
echo "== bench" ; timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== smoke" ; timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/smoke.log
echo "== ncu launch list (one timed bench step on a 16 s track)"
timeout 480 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode --track-seconds 16 > $OUT/launches_bench.log 2>&1 ; echo "rc=$?"
echo "== ncu full: al_istft inside the bench step (traffic)"
timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:'istft_pk[234]_kernel' -c 1 -o $OUT/prof_bench_istft -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode --track-seconds 60 > $OUT/prof_bench_istft.log 2>&1 ; echo "rc=$?"
ncu -i $OUT/prof_bench_istft.ncu-rep --page raw --csv > $OUT/prof_bench_istft_raw.csv 2>/dev/null
echo "== ncu full: spectral kernels, kernel_bench --once"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"stft_pk2_kernel|istft_pk2_kernel|ola_gather_kernel|resample_rb_kernel|istft_kernel|stft_kernel|gelu_bf16" \
    -o $OUT/prof_k -f python tools/kernel_bench.py --once --cases roformer_2048_441,htdemucs_4096_1024,mdx_6144_1024 --gelu > $OUT/prof_k.log 2>&1 ; echo "rc=$?"
ncu -i $OUT/prof_k.ncu-rep --page raw --csv > $OUT/prof_k_raw.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:istft_pk2 > $OUT/prof_k_istft_source.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:resample_rb > $OUT/prof_k_resample_source.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:'^(void )?(al::)?istft_kernel' > $OUT/prof_k_istft_generic_source.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:'^(void )?(al::)?stft_kernel' > $OUT/prof_k_stft_generic_source.csv 2>/dev/null
ncu -i $OUT/prof_k.ncu-rep --page source --csv -k regex:stft_pk2 > $OUT/prof_k_stft_pk_source.csv 2>/dev/null
# phase split of every captured kernel (SASS view cut at the CTA barriers)
for f in $OUT/prof_k_*_source.csv; do for i in 0 1 2 3 4 5; do python tools/ncu_phase_split.py $f $i 2>/dev/null; done > ${f%_source.csv}_phases.txt; done
rm -f $OUT/prof_k.ncu-rep
ls -la $OUT
