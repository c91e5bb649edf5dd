"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` export: totals per stall reason and the
hottest SASS instructions (with a little context), to see what a kernel's warps wait on."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
body = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
for r in body:
    for h in stall_cols:
        try:
            tot[h] += int(r[col[h]])
        except ValueError:
            pass
allsamp = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"instructions {len(body)}  samples {allsamp}")
for h, v in tot.most_common(12):
    print(f"  {h:28s} {v:8d}  {100.0 * v / max(allsamp, 1):5.1f}%")
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top_n]
print("hottest instructions (index: samples, executed, top stall, SASS):")
for i in sorted(order):
    r = body[i]
    st = max(stall_cols, key=lambda h: int(r[col[h]] or 0))
    print(f"  {i:5d}: {int(r[col['# Samples']]):6d} {r[col['Instructions Executed']]:>9s} {st[6:]:14s} {r[col['Source']].strip()[:110]}")
