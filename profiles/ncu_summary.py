#!/usr/bin/env python
"""Summarise one .ncu-rep (read with `ncu -i ... --page raw --csv`): the handful of counters
DESIGN.md and bench.py quote.  Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_i = hdr.index("Kernel Name")
for r in data:
    print("kernel:", r[name_i][:90])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"  {h:70s} {r[i]:>18s} {units[i]}")
    st = sorted(((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                 if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")), reverse=True)
    print("  stall reasons (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
