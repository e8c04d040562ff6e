#!/usr/bin/env python
"""Per device function: executed warp instructions, average active lanes and share of the idle lane slots.
usage: ncu_lost_lanes.py both.csv cuda.csv   (exports of tools/ncu_export.sh)"""
import sys, csv, collections, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_region_budget as R
full = {}
for r in csv.reader(open(sys.argv[2])):
    if len(r) == 2 and r[0].strip().isdigit():
        full[int(r[0])] = r[1]
fmap = R.enclosing_functions(full)
rows = list(csv.reader(open(sys.argv[1])))
cur = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        ii, it, isamp = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples"); continue
    if cur is None or len(r) < 8 or not r[0].strip().isdigit() or r[2] != '-':
        continue
    no = int(r[0]); fn = fmap.get(no, '?') if 'mos_device' in cur else cur
    a = agg[fn]; a[0] += int(r[ii] or 0); a[1] += int(r[it] or 0); a[2] += int(r[isamp] or 0)
tot = sum(a[0] for a in agg.values()); lost = sum(32 * a[0] - a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("total warp inst %.3e, avg lanes %.1f, idle lane slots %.3e" % (tot, sum(a[1] for a in agg.values()) / tot, lost))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print(f"{k:32s} warp-inst {a[0]/1e6:9.1f}M ({100*a[0]/tot:4.1f}%)  samples {100*a[2]/ts:4.1f}%  avg lanes {a[1]/max(a[0],1):5.1f}  idle share {100*(32*a[0]-a[1])/lost:5.1f}%")
