#!/usr/bin/env python
"""Print the handful of ncu metrics we read for a kernel: python tools/ncu_summary.py report.ncu-rep [launch index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
r = rows[2 + idx]
def g(n):
    return r[h.index(n)] if n in h else None
print("kernel", g("Kernel Name")[:90])
for n in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"):
    print("  %-70s %s" % (n, g(n)))
st = []
for i, n in enumerate(h):
    if "average_warps_issue_stalled" in n and n.endswith("_per_issue_active.ratio"):
        try:
            st.append((float(r[i]), n.split("issue_stalled_")[1].split("_per_issue")[0]))
        except ValueError:
            pass
print("  stall cycles per issued instruction:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:9]))
