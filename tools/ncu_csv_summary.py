"""Markdown summary of `ncu -i report --page raw --csv` exports (one kernel per file): the handful of metrics the
roofline discussion in DESIGN.md refers to.   python tools/ncu_csv_summary.py title raw.csv [raw2.csv ...] > out.md"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]


def main():
    print(f"# {sys.argv[1]}\n")
    for path in sys.argv[2:]:
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print(f"## {d['Kernel Name'][1]}\n\nsource: `{path}` (ncu --set full --clock-control none, one launch); grid "
              f"{d['Grid Size'][1]}, block {d['Block Size'][1]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k][1]} | {d[k][0]} |")
        print()


if __name__ == "__main__":
    main()
