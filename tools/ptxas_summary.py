#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs (feltor_b200/csrc/build/*.ptxas.log): registers, spills, smem per kernel."""
import re, subprocess, sys, glob, os
logs = sys.argv[1:] or glob.glob(os.path.join(os.path.dirname(__file__), "..", "feltor_b200", "csrc", "build", "*.ptxas.log"))
rows = []
for log in logs:
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n\s*ptxas info\s*:\s*Used (\d+) registers(.*)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void dgb::", "")
        smem = re.search(r"(\d+) bytes smem", m.group(6))
        rows.append((name, int(m.group(5)), int(m.group(2)), int(m.group(3)), int(smem.group(1)) if smem else 0))
flt = os.environ.get("FILTER")
for r in sorted(rows):
    if flt and not re.search(flt, r[0]): continue
    print(f"{r[0][:90]:90s} regs={r[1]:3d} stack={r[2]:4d} spill_st={r[3]:4d} smem={r[4]}")
