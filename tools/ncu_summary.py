#!/usr/bin/env python
"""One-screen summary of an .ncu-rep (raw page): duration, DRAM traffic, pipe utilisation, occupancy, stalls."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
I = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("==", r[I["Kernel Name"]][:100], "grid", r[I["Grid Size"]], "block", r[I["Block Size"]])
    for k in keys:
        if k in I:
            print(f"   {k:70s} {r[I[k]]:>14s} {units[I[k]]}")
    st = [(h, float(r[I[h]] or 0)) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st.sort(key=lambda kv: -kv[1])
    print("   stalls/issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for h, v in st[:7]))
