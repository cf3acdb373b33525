"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python profiles/summarize_launches.py profiles/r01_launches_v1.csv"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:80]
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us of kernel time (cold-cache, serialised)")
print(f"{'us':>10s} {'n':>5s} {'share':>6s}  kernel")
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {name}")
