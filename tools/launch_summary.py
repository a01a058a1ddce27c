"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE training step
(the launches between the last two embed_fwd kernels).  Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
i0 = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i0:]))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "embed_fwd" in n]
a, b = idx[-2], idx[-1]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[a:b]:
    n = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
    agg[n][0] += 1
    agg[n][1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"one step: {b - a} launches, {tot:.1f} us summed kernel time (ncu: serialised, cold caches)")
for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:4d} {100 * v[1] / tot:5.1f}% {n}")
