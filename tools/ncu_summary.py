#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of metrics DESIGN.md and
bench.py quote.  Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb_per_issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            print("== %s :: %s" % (path.split("/")[-1], name))
            for metric, short in WANT:
                if metric in hdr:
                    i = hdr.index(metric)
                    print("   %-24s %s %s" % (short, r[i], units[i]))


if __name__ == "__main__":
    main()
