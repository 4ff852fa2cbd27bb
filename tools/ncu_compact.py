#!/usr/bin/env python
"""Compact per-kernel table from an `ncu --metrics ... --csv --log-file` capture."""
import collections
import csv
import sys

SHORT = collections.OrderedDict([
    ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'dramR_MB'), ('dram__bytes_write.sum', 'dramW_MB'),
    ('lts__t_bytes.sum', 'L2_MB'), ('sm__inst_executed.sum', 'Minst'), ('sm__inst_executed_pipe_fma.sum', 'Mfma'),
    ('sm__inst_executed_pipe_alu.sum', 'Malu'), ('sm__inst_executed_pipe_lsu.sum', 'Mlsu'),
    ('sm__inst_issued.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'Mbankconf'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'Msmem_wf'), ('sm__cycles_elapsed.max', 'kcyc'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('launch__registers_per_thread', 'regs'),
])
SCALE = {'us': 1e-3, 'dramR_MB': 1e-6, 'dramW_MB': 1e-6, 'L2_MB': 1e-6, 'Minst': 1e-6, 'Mfma': 1e-6, 'Malu': 1e-6,
         'Mlsu': 1e-6, 'Mbankconf': 1e-6, 'Msmem_wf': 1e-6, 'kcyc': 1e-3}


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    by = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        key = (r[idx['ID']], r[idx['Kernel Name']][:44])
        try:
            v = float(r[idx['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        by.setdefault(key, {})[r[idx['Metric Name']]] = v
    for k, m in by.items():
        print(f"{k[0]:>3} {k[1]:44s} " + ' '.join(
            f"{s}={m[n] * SCALE.get(s, 1):.4g}" for n, s in SHORT.items() if n in m))


if __name__ == '__main__':
    main(sys.argv[1])
