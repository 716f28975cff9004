#!/usr/bin/env python
"""Summarise an ncu report (--set full) into a small text file for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/TAG/prof.ncu-rep profiles/NAME.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % rep)
        for r in rows[2:]:
            f.write("\n== %s  (launch id %s)\n" % (r[hdr.index("Kernel Name")], r[hdr.index("ID")]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("  %-85s %s %s\n" % (k, r[i], units[i]))
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
                u = units[hdr.index("dram__bytes_read.sum")]
                f.write("  %-85s %.3f %s\n" % ("traffic = dram read + write", rd + wr, u))
            except Exception:
                pass
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
