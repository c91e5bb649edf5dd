"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time and share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.3f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare SHARES, not absolutes)")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:9.3f} ms {100 * v / T:5.1f}%  n={cnt[k]:4d}  {k[:120]}")
