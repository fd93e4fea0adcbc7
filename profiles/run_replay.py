#!/usr/bin/env python
"""Profiling driver: N replays of the recorded DW trajectory of a config through the C ABI
(device-resident duals), nothing else.  Used under ncu (see profiles/README.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pis, phases, _ = load_trajectory(tag)
prob = P.make_config(tag)
eng = Engine(prob, fetch_x=False)
pis_dev = torch.from_numpy(pis).cuda()
for rep in range(reps):
    eng.reset()
    eng.initial_solve_device()
    if rep == 0:
        it = eng.block_iterations()
        print("cold iterations/block: mean %.1f max %d p90 %d" % (it.mean(), it.max(), np.percentile(it, 90)))
    for r in range(len(phases)):
        eng.iterate_device(pis_dev[r].data_ptr(), phases[r])
        if rep == 0:
            it = eng.block_iterations()
            print("round %d iterations/block: mean %.2f max %d zero-blocks %d" % (r, it.mean(), it.max(), (it == 0).sum()))
    eng.sync()
c = eng.counters()
print("replays", reps, c)
eng.close()
