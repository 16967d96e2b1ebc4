#!/usr/bin/env python
"""Instruction mix and stall summary of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K`.
   python tools/ncu_mix.py file.csv   (first kernel instance in the file)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# the file holds one block per kernel instance: "Kernel Name" line, header line, instruction lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = blocks[which]
h = b["hdr"]
ix = {k: i for i, k in enumerate(h)}
print(b["name"][:110])
mix, stalls = collections.Counter(), collections.Counter()
total = 0
samples_by_op = collections.Counter()
for r in b["rows"]:
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    n = int(float(r[ix["Instructions Executed"]] or 0))
    mix[op] += n
    total += n
    samples_by_op[op] += int(float(r[ix["# Samples"]] or 0))
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            stalls[k] += int(float(r[ix[k]] or 0))
print("total warp instructions:", total)
for op, n in mix.most_common(28):
    print(f"  {op:12s} {n:10d} {100.*n/total:5.1f}%   samples {samples_by_op[op]}")
ts = sum(stalls.values())
print("stall samples (all):", ts)
for k, n in stalls.most_common(12):
    print(f"  {k:28s} {n:8d} {100.*n/ts:5.1f}%")
