#!/usr/bin/env python
"""Time-weighted tensor-pipe activity per tcgen05 kernel from an
`ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_{active,elapsed},sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum --csv` log."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ii, ki, ni, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
per = {}
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ni] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1e-3)
    per.setdefault(r[ii], {"k": r[ki]})[r[ni]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for e in per.values():
    name = re.sub(r"\(.*", "", e["k"].replace("void ", "").replace("<unnamed>::", ""))
    t = e.get("gpu__time_duration.sum", 0.0)
    a = agg[name]
    a[0] += 1; a[1] += t
    a[2] += t * e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a[3] += t * e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    a[4] += e.get("sm__inst_executed_pipe_tensor.sum", 0.0)
print("tensor-pipe counters of every tcgen05 launch (ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_{active,elapsed},sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum)")
print("of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph` on one B200; time-weighted means per kernel")
print(f"{'kernel':44s} {'launches':>8s} {'avg us':>8s} {'tensor pipe active % (of active cycles)':>40s} {'(of elapsed)':>14s} {'tensor instr / launch':>22s}")
for name, (n, t, wa, we, ti) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name:44s} {n:8d} {t / n:8.1f} {wa / t:40.1f} {we / t:14.1f} {ti / n:22.0f}")
