#!/usr/bin/env python
"""Path 5 (packed shared-memory tableau) on a config, launch by launch: kernel time of every round of two replays of the
recorded DW trajectory, and the hand-over / re-pack counters (Engine.sparse_stats).  DWGPU_SPARSE=0 gives the path-3 numbers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "B"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pis, phases, _ = load_trajectory(tag)
prob = P.make_config(tag)
eng = Engine(prob, fetch_x=False)
paths = eng.block_paths()
print("paths", {int(p): int((paths == p).sum()) for p in set(paths)})
pis_dev = torch.from_numpy(pis).cuda()
for rep in range(reps):
    eng.reset()
    eng.sync()
    ms = []
    eng.initial_solve_device()
    eng.sync()
    ms.append(eng.last_kernel_ms()[1])
    st = [eng.sparse_stats()]
    for r in range(len(phases)):
        eng.iterate_device(pis_dev[r].data_ptr(), phases[r])
        eng.sync()
        ms.append(eng.last_kernel_ms()[1])
        st.append(eng.sparse_stats())
    print("replay %d: total kernel ms %.3f" % (rep, sum(ms)))
    print("  ms per launch:", " ".join("%.3f" % v for v in ms))
    keys = ["to_large_heap", "to_dense_heap", "to_dense_residual", "repacks", "batched_sweeps", "union_checks", "probe_finished", "probe_passed_on"]
    prev = {k: 0 for k in keys} if rep == 0 else last
    for r, s in enumerate(st):
        d = {k: s[k] - prev[k] for k in keys}
        if r < 8 or any(d[k] for k in keys[:5]):
            print("  launch %2d:" % r, " ".join("%s=%d" % (k, d[k]) for k in keys))
        prev = s
    last = st[-1]
    print("  heaps", st[-1]["heap_small"], st[-1]["heap_large"])
c = eng.counters()
print("replays", reps, c)
eng.close()
