#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of metrics DESIGN.md / bench.py cite.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block ", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# {rep}")
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")]
    print(f"\n## {name}")
    for h, u, v in zip(hdr, units, vals):
        if any(k.strip() == h or (k.endswith(" ") and h == k.strip()) for k in KEYS):
            print(f"{h:95s} {u:16s} {v}")
# SASS opcode mix (needs --import-source / sass in the report)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
if len(srows) > 3:
    h = srows[1]
    if "Source" in h and "Instructions Executed" in h:
        isrc, iex = h.index("Source"), h.index("Instructions Executed")
        mix = {}
        for r in srows[2:]:
            try:
                op = r[isrc].split()[0] if not r[isrc].lstrip().startswith("@") else r[isrc].split()[1]
                mix[op.split(".")[0]] = mix.get(op.split(".")[0], 0) + int(r[iex])
            except Exception:
                pass
        tot = sum(mix.values()) or 1
        print("\n## executed warp-instructions by opcode (first kernel in the report)")
        for op, n in sorted(mix.items(), key=lambda x: -x[1])[:16]:
            print(f"{op:14s} {n:14d} {100.0 * n / tot:6.2f} %")
