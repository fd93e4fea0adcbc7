#!/usr/bin/env python
"""Where a CTA's time goes: replays a config's recorded DW trajectory with a -DDWK_TIMERS=1 build of the engine
(DWGPU_LIB=alt/libdwgpu_timers.so) and prints, per round and in total, the share of CTA cycles per kernel phase.
Thread 0 of every CTA charges clock64() differences to the current phase (simplex_smem.cuh: phase())."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import engine as E  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402

NAMES = ["prologue (state in)", "feasibility / phase-1 prices", "reduced-cost row", "pricing argmax", "column fetch + list", "ratio test (warp 0)", "capacity check / re-pack", "sweep", "x extraction + residual", "batched sweep", "epilogue", "bookkeeping + pivot row", "-", "-", "-", "CTAs"]

tag = sys.argv[1] if len(sys.argv) > 1 else "B"
pis, phases, _ = load_trajectory(tag)
prob = P.make_config(tag)
eng = E.Engine(prob, fetch_x=False)
lib = E.load_library()
fn = lib.dwgpu_debug_phases
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()


def grab():
    assert fn(buf, 1) == 0
    return np.array(list(buf), dtype=np.float64)


pis_dev = torch.from_numpy(pis).cuda()
eng.reset()
grab()
total = np.zeros(16)
rows = []
eng.initial_solve_device()
eng.sync()
v = grab(); total += v; rows.append(("cold", v, sum(eng.last_kernel_ms())))
for r in range(len(phases)):
    eng.iterate_device(pis_dev[r].data_ptr(), phases[r])
    eng.sync()
    v = grab(); total += v; rows.append(("round %d" % r, v, sum(eng.last_kernel_ms())))
print("%-10s %9s %12s  " % ("launch", "ms", "Mcycles") + "  ".join("%2d" % i for i in range(12)))
for name, v, ms in rows:
    cyc = v[:12].sum()
    if cyc == 0:
        continue
    print("%-10s %9.3f %12.1f  " % (name, ms, cyc / 1e6) + "  ".join("%2.0f" % (100 * x / cyc) for x in v[:12]))
cyc = total[:12].sum()
print("\ntotal CTA cycles %.3e over %d CTAs" % (cyc, total[15]))
for i in range(12):
    print("  %2d %-28s %6.2f %%" % (i, NAMES[i], 100 * total[i] / cyc))
if total[13] > 0:
    print("pivot rows fetched inside the sweep: %.0f items of 32 B, %.2f non-zero doubles per item (4 = perfectly packed)" % (total[13], total[12] / total[13]))
eng.close()
