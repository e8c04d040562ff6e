#!/usr/bin/env python
"""Aggregate the ncu source page (SASS view) by opcode: executed warp instructions and stall samples."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter(); total = 0; tsamp = 0
top = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else "?"
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["# Samples"]] or 0)
    ex[op] += n; st[op] += s; total += n; tsamp += s
    top.append((s, n, r[ci["Address"]], src.strip()[:90]))
print(f"total warp-inst {total:,}  samples {tsamp:,}")
print("opcode       inst      %inst  %samples")
for op, n in ex.most_common(28):
    print(f"{op:10s} {n:12,d} {100*n/total:6.1f} {100*st[op]/max(tsamp,1):6.1f}")
if len(sys.argv) > 2:
    print("\nhottest instructions by samples")
    for s, n, a, src in sorted(top, reverse=True)[:int(sys.argv[2])]:
        print(f"{s:7d} {n:10d} {a} {src}")
