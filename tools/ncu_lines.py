#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source sass,cuda --csv` output per CUDA source line.
usage: ncu -i rep --page source --print-source sass,cuda --csv > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv
import collections
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, ''])
H = None
cur = None
fn = ''
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        H = r
        ie, ss = H.index('Instructions Executed'), H.index('# Samples')
        continue
    if len(r) == 2 and r[0] == 'Function Name':
        fn = r[1][:60]
        continue
    if H is None or len(r) <= ie:
        continue
    if r[0].strip().isdigit():
        cur = (fn, int(r[0]))
        agg[cur][2] = r[1]
    if not r[2].strip():
        continue
    try:
        v, s = int(r[ie]), int(r[ss])
    except ValueError:
        continue
    agg[cur][0] += v
    agg[cur][1] += s
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values()) or 1
print("total warp instructions", tot, "samples", tots)
for (f, l), (v, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v/tot*100:5.1f}% instr  {s/tots*100:5.1f}% smp  L{l}: {t.strip()[:105]}")
