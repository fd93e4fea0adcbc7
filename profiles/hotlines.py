#!/usr/bin/env python
"""Rank CUDA source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: python profiles/hotlines.py <csv> [top N]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, idx = None, None, None
agg = defaultdict(lambda: [0, defaultdict(int), ""])
for r in csv.reader(open(path, errors="replace")):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None:
        continue
    try:
        ns = int(r[idx["# Samples"]])
    except (ValueError, IndexError):
        continue
    line = r[0]
    if r[2] == "-":                      # source-line row: remember the text
        agg[(cur_file, line)][2] = r[1].strip()
        cur_line = line
        continue
    key = (cur_file, line if line not in ("", "-") else cur_line)
    a = agg[key]
    a[0] += ns
    for h, i in idx.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                a[1][h[6:]] += int(r[i])
            except ValueError:
                pass
tot = sum(a[0] for a in agg.values())
print(f"total samples {tot}")
by_reason = defaultdict(int)
for a in agg.values():
    for k, v in a[1].items():
        by_reason[k] += v
print("by reason:", ", ".join(f"{k} {100*v/max(tot,1):.1f}%" for k, v in sorted(by_reason.items(), key=lambda t: -t[1])[:8]))
for (f, line), a in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in sorted(a[1].items(), key=lambda t: -t[1])[:3] if v)
    print(f"{a[0]:8d} {100*a[0]/max(tot,1):5.1f}%  {f}:{line:>4}  {a[2][:90]}   [{st}]")
