#!/usr/bin/env python
"""Static code footprint per device function next to how often it runs (from `ncu_export.sh` output): which functions make up the
instruction working set of a kernel.  usage: ncu_code_footprint.py both.csv cuda.csv [warm_fraction=0.001]
A SASS instruction counts as warm when it executes at least warm_fraction x the most executed instruction's count."""
import collections
import csv
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_region_budget as R

full = {}
for r in csv.reader(open(sys.argv[2])):
    if len(r) == 2 and r[0].strip().isdigit():
        full[int(r[0])] = r[1]
fmap = R.enclosing_functions(full)
warm_frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
cur, curline, i_inst, recs = None, None, None, []
for r in csv.reader(open(sys.argv[1])):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        i_inst = r.index("Instructions Executed")
        continue
    if cur is None or len(r) < 8:
        continue
    if r[0].strip().isdigit() and r[2] == "-":
        curline = (cur, int(r[0]))
        continue
    if r[0] == "" and r[2].startswith("0x") and curline is not None:
        f, line = curline
        fn = fmap.get(line, "?") if "mos_device" in f else os.path.basename(f)
        recs.append((fn, int(r[i_inst] or 0)))
mx = max(e for _, e in recs)
stat = collections.defaultdict(lambda: [0, 0, 0])
for fn, e in recs:
    s = stat[fn]
    s[0] += 1
    s[1] += e
    s[2] += e >= warm_frac * mx
tot = sum(s[1] for s in stat.values())
print(f"SASS instructions {len(recs)} ({len(recs) * 16 / 1024:.1f} KB), warm (>= {warm_frac:g} x max) {sum(s[2] for s in stat.values())} "
      f"({sum(s[2] for s in stat.values()) * 16 / 1024:.1f} KB)")
print(f"{'function':34s} {'static':>6s} {'warm':>6s} {'% of executed':>13s}")
for fn, s in sorted(stat.items(), key=lambda kv: -kv[1][2])[:40]:
    print(f"{fn:34s} {s[0]:6d} {s[2]:6d} {100 * s[1] / tot:13.2f}")
