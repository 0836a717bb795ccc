#!/usr/bin/env python
"""One step of the hot path out of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_step.py launches.csv [which]  -> the kernels between two consecutive bm25_plan_terms launches."""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [(r[ki][:72], float(r[vi].replace(",", "")) / (1000.0 if r[ui] in ("ns", "nsecond") else 1.0)) for r in rows[1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(data) if "bm25_plan_terms" in k]
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(idx) // 2
a, b = idx[which], idx[which + 1] if which + 1 < len(idx) else len(data)
tot = 0.0
for k, v in data[a:b]:
    print(f"{v:10.1f} us  {k}")
    tot += v
print(f"{tot:10.1f} us  total")
