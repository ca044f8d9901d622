#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-launch key metrics as a small CSV/markdown.
Usage: python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep > profiles/rNN/prof_summary.md"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}` ({len(data)} launches; `--set full --clock-control none`)\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join(r[col["Kernel Name"]][:60] for r in data) + " |")
    for k in KEYS:
        if k in col:
            print(f"| {k} | {units[col[k]]} | " + " | ".join(r[col[k]] for r in data) + " |")
if __name__ == "__main__":
    main(sys.argv[1])
