#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_shares.py file.csv [skip]"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
agg = OrderedDict()
n = 0
for r in rows[1:]:
    if int(r[iid]) < skip:
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void\s+", "", name).replace("btsb::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot / 1e3:.1f} us total (cold-cache, serialised under ncu: compare SHARES)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t / tot * 100:5.1f}%  {c:4d} x {t / c / 1e3:9.1f} us  {k}")
