#!/usr/bin/env python3
"""Summarise an .ncu-rep (first profiled kernel): python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [n_warp_steps]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return m.get(k, ("n/a", ""))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__warps_eligible.avg.per_cycle_active"]
for k in keys:
    v, u = g(k)
    print(f"{k:75s} {v} {u}")
print("-- stall cycles per issued instruction --")
st = [(float(v[0].replace(',', '')), h.split('issue_stalled_')[1].split('_per_issue')[0]) for h, v in m.items()
      if 'smsp__average_warps_issue_stalled_' in h and h.endswith('_per_issue_active.ratio') and v[0] not in ('', 'n/a')]
for val, name in sorted(st, reverse=True):
    if val > 0.005:
        print(f"   {name:28s} {val:.3f}")
if len(sys.argv) > 2:
    n = float(sys.argv[2])
    tot = float(g("smsp__inst_executed.sum")[0].replace(',', ''))
    print(f"instructions per warp-step: {tot / n:.1f}")
