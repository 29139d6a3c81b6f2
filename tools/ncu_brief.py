#!/usr/bin/env python
"""Print the handful of figures that decide what bounds a kernel from an ncu report (first kernel in it).
usage: python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {k: (v[i], u[i]) for i, k in enumerate(h)}
for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
          "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
          "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"):
    if k in d:
        print(f"{k:75s} {d[k][0]} {d[k][1]}")
try:
    ipc = float(d["smsp__inst_executed.sum"][0].replace(",", "")) / (float(d["smsp__cycles_active.avg"][0].replace(",", "")) * 592)
    print(f"{'issue slots used (inst / (592 SMSP x active cycles))':75s} {ipc:.3f}")
except Exception:
    pass
st = []
for k, (x, _) in d.items():
    if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        try:
            st.append((float(x), k.split("issue_stalled_")[1].split("_per_issue")[0]))
        except ValueError:
            pass
print("stall cycles per issue:", ", ".join(f"{n} {a:.2f}" for a, n in sorted(st, reverse=True)[:7]))
