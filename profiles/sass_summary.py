#!/usr/bin/env python
"""Regenerates profiles/sass_summary.md from `cuobjdump -sass` of the built engine (runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "dwsolver_b200", "lib", "libdwgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
keys = ["UBLKCP", "SYNCS", "UCGABAR", "DMMA", "DFMA", "CREDUX", "LDG.E.128", "STG.E.128", "ATOMS", "BAR.SYNC", "LDL/STL", "CCTL/PREFETCH"]
rows, cur, it = [], None, iter(names)
for line in sass.split("\n"):
    if "Function :" in line:
        cur = collections.Counter(); rows.append((next(it).split("(")[0].replace("void ", ""), cur)); continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for k in keys[:-2]:
        if op.startswith(k) or (k in ("LDG.E.128", "STG.E.128") and op.startswith(k[:3]) and ".128" in op):
            cur[k] += 1
    if op.startswith("LDL") or op.startswith("STL"):
        cur["LDL/STL"] += 1
    if op.startswith("CCTL") or "PREFETCH" in op:
        cur["CCTL/PREFETCH"] += 1
rows.sort(key=lambda r: -sum(r[1].values()))
out = ["# SASS summary of dwsolver_b200/lib/libdwgpu.so (sm_100a), `cuobjdump -sass`, round 2 (final build)", "",
       "Counts of the mnemonics that identify the Blackwell-era paths (B200_PROFILING.md): `UBLKCP` = TMA bulk copy (`cp.async.bulk`),",
       "`SYNCS` = mbarrier, `UCGABAR` = cluster barrier, `DMMA` = FP64 tensor-core MMA (`mma.sync.m8n8k4.f64`), `CREDUX` = warp reduction",
       "(`redux.sync`).  `UTMALDG` (tensor-map TMA) does not occur: every bulk transfer here is a contiguous image, not a tiled tensor;",
       "`LDL/STL` = local-memory (spill) instructions.  Regenerate: `python profiles/sass_summary.py > profiles/sass_summary.md`.", "",
       "| kernel | " + " | ".join(keys) + " |", "|---|" + "---:|" * len(keys)]
for name, c in rows:
    if sum(c.values()) == 0:
        continue
    out.append(f"| `{name}` | " + " | ".join(str(c[k]) for k in keys) + " |")
sys.stdout.write("\n".join(out) + "\n")
