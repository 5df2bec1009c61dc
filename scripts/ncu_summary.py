#!/usr/bin/env python
"""Summarise an ncu report (raw page exported as CSV) into one row per launch.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv [out.csv]"""
import csv
import sys

KEEP = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_sel"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_no_inst"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n) for (k, n) in KEEP if k in hdr]
    out = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    out.writerow([n + ("[%s]" % units[i] if units[i] else "") for (i, n) in cols])
    for r in rows[2:]:
        vals = []
        for (i, n) in cols:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("fcd::", "")[:48]
            vals.append(v)
        out.writerow(vals)


if __name__ == "__main__":
    main()
