#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples (from `ncu --page source --print-source cuda,sass --csv`)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cur_file = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 8 and r[0].strip().isdigit():
        try:
            agg.append((int(r[7] or 0), int(r[6] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot_i = sum(a[0] for a in agg); tot_s = sum(a[1] for a in agg)
print(f"total inst {tot_i:,} samples {tot_s:,}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = 0 if len(sys.argv) <= 3 else 1
for a in sorted(agg, key=lambda x: -x[key])[:n]:
    print(f"{100*a[0]/tot_i:5.1f}%i {100*a[1]/tot_s:5.1f}%s {a[2]}:{a[3]}  {a[4]}")
