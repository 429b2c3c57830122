#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: executed-instruction mix by opcode, stall reasons, hottest lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(ends) > 1:                      # several launches in one report: summarise the first
    rows = rows[:ends[1]]
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); tot = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[col["Source"]].strip()
    tok = sass.split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    n = int(r[col["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            stalls[k] += int(r[col[k]] or 0)
    recs.append((int(r[col["# Samples"]] or 0), n, sass))
print("total warp-instructions executed:", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"  {n:12d} {100*n/tot:5.1f}%  {op}")
st = sum(stalls.values())
print("stall samples:")
for k, n in stalls.most_common(12):
    print(f"  {n:8d} {100*n/st:5.1f}%  {k}")
print("hottest instructions by samples:")
for s, n, sass in sorted(recs, reverse=True)[:25]:
    print(f"  {s:6d} {n:10d}  {sass[:100]}")
