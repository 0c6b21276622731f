"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py <launches.csv> [steps_in_capture]"""
import collections
import csv
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for x in rows:
    a = agg.setdefault(x["Kernel Name"][:90], [0, 0.0])
    a[0] += 1
    a[1] += float(x["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total, {tot / 1e3 / steps:.1f} us per step over {steps} steps (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3 / steps:9.1f} us/step {v[0] / steps:6.1f} x/step {v[1] / v[0] / 1e3:8.2f} us/launch {100 * v[1] / tot:5.1f}%  {k}")
