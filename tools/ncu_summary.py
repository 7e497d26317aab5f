#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the text summary kept under profiles/.
usage: tools/ncu_summary.py report.ncu-rep out.txt "header comment" """
import csv
import subprocess
import sys

rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
stalls = sorted([(float(v), h) for h, (u, v) in d.items()
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v], reverse=True)
lines = ["# " + header]
for k in keys:
    if k in d:
        lines.append("%-75s %s %s" % (k, d[k][1], d[k][0]))
lines.append("# warp stall reasons (warps per issue-active cycle), descending")
for v, h in stalls[:10]:
    lines.append("%-75s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:26]))
