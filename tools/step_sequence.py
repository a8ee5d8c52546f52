#!/usr/bin/env python
"""Print the kernels of ONE step (between the last two smooth_noise launches) from an ncu launch list, in launch order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
seq = []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    n = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    seq.append((n, float(r[vi].replace(",", "")) / 1e3, r[gi]))
idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith("smooth_noise")]
s, e = idx[-2], idx[-1]
tot = 0.0
agg = {}
for n, t, g in seq[s:e]:
    tot += t
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
    if len(sys.argv) > 2:
        print(f"{n[:44]:44s} {t:8.1f} {g}")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:60]:60s} {c:4d} {t:9.1f} us {t / c:7.1f} avg {100 * t / tot:5.1f}%")
print(f"one step: {tot:.1f} us over {e - s} launches")
