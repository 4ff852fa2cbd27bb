#!/bin/bash
# Quick GPU call: kernel bench + ncu metric CSV for the spectral kernels (small outputs only).
# usage: bash tools/gpu_quick.sh TAG [kernel-regex-for-full-capture [kernel_bench case]]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -x tools/ubench/f32x2 ]; then tools/ubench/f32x2 > $OUT/f32x2.txt 2>&1; cat $OUT/f32x2.txt; fi
timeout 300 python tools/kernel_bench.py > $OUT/kernel_bench.jsonl 2> $OUT/kernel_bench.err ; echo "kernel_bench rc=$?" ; cat $OUT/kernel_bench.jsonl ; tail -3 $OUT/kernel_bench.err
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fp32.sum
timeout 600 ncu --metrics $METRICS --clock-control none -k regex:'stft_kernel|istft_kernel|ola_gather_kernel|resample_kernel' \
    --csv --log-file $OUT/ncu_metrics.csv python tools/kernel_bench.py --once > $OUT/ncu_metrics.log 2>&1 ; echo "ncu rc=$?"
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -c 1 -o $OUT/prof -f python tools/kernel_bench.py --once --cases ${3:-roformer_2048_441} > $OUT/prof.log 2>&1 ; echo "ncu full rc=$?"
  ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
  ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
  ncu -i $OUT/prof.ncu-rep --page details > $OUT/prof_details.txt 2>/dev/null
  ls -la $OUT/prof.ncu-rep
  [ $(stat -c %s $OUT/prof.ncu-rep) -gt 30000000 ] && rm -f $OUT/prof.ncu-rep
fi
du -sh $OUT
