#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
skip = set(sys.argv[2:])
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    if name in skip: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={v[0]:4d} total={v[1] / 1e3:10.1f} us avg={v[1] / v[0] / 1e3:8.1f} us share={v[1] / tot:.3f}")
