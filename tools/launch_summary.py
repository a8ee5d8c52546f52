#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = r[ki].replace("<unnamed>::", "").replace("void ", "")[:70]
    d[n][0] += 1
    d[n][1] += v
tot = sum(v[1] for v in d.values())
print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for n, (c, t) in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f"{n:70s} {c:5d} {t / 1e3:10.1f} {t / c / 1e3:8.1f} {100 * t / tot:5.1f}%")
print(f"total {tot / 1e3:.1f} us over {sum(v[0] for v in d.values())} launches")
