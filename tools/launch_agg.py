"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_agg.py file.csv [top]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = re.sub(r"\(.*", "", r[ki])
    n = re.sub(r"^void ", "", n)[:90]
    agg[n][0] += 1
    agg[n][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t / 1e3:10.1f} us {c:5d} {t / tot * 100:5.1f}%  {n}")
