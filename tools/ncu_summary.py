"""Compact per-launch summary of an `ncu --page raw --csv` export (the raw file has ~1000 columns):
   python tools/ncu_summary.py gpurun_out/x_raw.csv > profiles/x_summary.csv"""
import csv
import sys

COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "duration_us"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__inst_executed.sum", "warp_insts"), ("smsp__issue_active.avg.pct", "issue_active_pct"),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    have = [(c, n) for c, n in COLS if c in idx]
    w = csv.writer(sys.stdout)
    w.writerow([n + ("[%s]" % units[idx[c]] if units[idx[c]] else "") for c, n in have])
    for r in data:
        w.writerow([(r[idx[c]].split("(")[0].replace("<unnamed>::", "").replace("void ", "") if n == "kernel" else r[idx[c]]) for c, n in have])


if __name__ == "__main__":
    main()
