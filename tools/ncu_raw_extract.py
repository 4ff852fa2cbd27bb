#!/usr/bin/env python
"""Print the key metrics + stall-reason samples of each kernel in an `ncu --page raw --csv` export."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', vals[hdr.index('Kernel Name')], 'grid', vals[hdr.index('Grid Size')], 'block', vals[hdr.index('Block Size')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:78s} {vals[i]:>16s} {units[i]}')
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
                try:
                    stalls.append((float(vals[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1
        print('  stall samples: ' + ', '.join(f'{n}={100 * v / tot:.1f}%' for v, n in sorted(stalls, reverse=True)[:9]))


if __name__ == '__main__':
    main(sys.argv[1])
