#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.  usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    print("=====", r[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx: print("  %-62s %s %s" % (w, r[idx[w]], rows[1][idx[w]]))
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try: st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("  stalls (warps per issue):", ", ".join("%s=%.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))
