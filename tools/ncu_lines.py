#!/usr/bin/env python
"""Per-source-line totals from `ncu --page source --csv --print-source sass,cuda` output:
python tools/ncu_lines.py dump.csv [file-substring] -> line, instructions executed, stall samples, smem wavefronts"""
import csv
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(open(path)))
cur_file = ""
hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    if want and want not in cur_file:
        continue
    d = dict(zip(hdr, r))
    try:
        out.append((cur_file.split("/")[-1], int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]),
                    int(d["L1 Wavefronts Shared"]), int(d["stall_long_sb"]), int(d["stall_barrier"]), int(d["stall_short_sb"])))
    except ValueError:
        pass
tot_i = sum(o[3] for o in out) or 1
tot_s = sum(o[4] for o in out) or 1
print(f"total inst {tot_i} samples {tot_s}")
print("file:line | inst% | samples% | smem wavefronts | long_sb | barrier | short_sb | source")
for o in sorted(out, key=lambda o: -o[4])[:40]:
    print(f"{o[0]}:{o[1]} | {100 * o[3] / tot_i:.1f} | {100 * o[4] / tot_s:.1f} | {o[5]} | {o[6]} | {o[7]} | {o[8]} | {o[2]}")
