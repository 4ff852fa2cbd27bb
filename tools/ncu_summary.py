"""Compact per-kernel summary of an `ncu --page raw --csv` export: the metrics B200_PROFILING.md names."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst_%"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_%"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleeping"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("#", r[idx["Kernel Name"]][:110])
        out = []
        for key, name in WANT:
            if key in idx and r[idx[key]] != "":
                out.append(f"{name}={r[idx[key]]}{(' ' + units[idx[key]]) if units[idx[key]] not in ('', '%') else ''}")
        print("  " + "; ".join(out))


if __name__ == "__main__":
    main()
