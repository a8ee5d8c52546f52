#!/usr/bin/env python
"""DRAM traffic of the tensor-core GEMM launches of ONE training step, from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc_kernel --csv` log.
Writes profiles/gemm_traffic.json (read by bench.py for roofline.traffic)."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ii, ni, vi, ui = h.index("ID"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
per = {}
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1}.get(u, 1)
    per.setdefault(r[ii], {})[r[ni]] = v * scale
ids = sorted(per, key=int)
n_step = int(sys.argv[3]) if len(sys.argv) > 3 else 68
last = ids[-n_step:]                       # the last full step's launches
rd = sum(per[i].get("dram__bytes_read.sum", 0) for i in last)
wr = sum(per[i].get("dram__bytes_write.sum", 0) for i in last)
ns = sum(per[i].get("gpu__time_duration.sum", 0) for i in last)
out = {"launches": len(last), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd + wr) / len(last),
       "time_us": ns / 1e3, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the gemm_tc_kernel launches of one step"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
