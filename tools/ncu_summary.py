#!/usr/bin/env python
"""Summarises an .ncu-rep (read on the CPU box): headline metrics, stall reasons and the hottest SASS lines."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__inst_executed.sum.per_cycle_active", "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
              "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]:
        if k in d:
            print(f"   {k:75s} {d[k]:>16s} {u[k]}")
    stalls = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(v), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("   stalls:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print(f"-- source: {len(data)} SASS instructions, {tot} samples; top {top_n}:")
keys = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_math", "stall_lg", "stall_mio", "stall_not_selected"]
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:top_n]:
    extra = " ".join(f"{k[6:]}={r[idx[k]]}" for k in keys if k in idx and r[idx[k]] not in ("0", ""))
    print(f"{int(r[idx['# Samples']]):7d} {100*int(r[idx['# Samples']])/max(tot,1):5.1f}%  {r[idx['Source']][:64]:64s} {extra}")
