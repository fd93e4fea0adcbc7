#!/usr/bin/env python
"""How does the per-block solve time of config B change with the number of blocks on the GPU
(working set / TLB reach / waves)?  Replays the first rounds of the recorded trajectory on the first
K blocks and prints the kernel time per launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "B"
pis, phases, _ = load_trajectory(tag)
pis_dev = torch.from_numpy(pis).cuda()
R = min(6, len(phases))
for K in [int(a) for a in sys.argv[2:]] or [592, 1184, 2368, 4736, 8192]:
    prob = P.make_config(tag, block_range=range(K))
    eng = Engine(prob, fetch_x=False)
    for rep in range(2):
        eng.reset()
        ms = []
        c0 = eng.counters()
        eng.initial_solve_device(); ms.append(max(eng.last_kernel_ms()))
        for r in range(R):
            eng.iterate_device(pis_dev[r].data_ptr(), phases[r]); ms.append(max(eng.last_kernel_ms()))
        c1 = eng.counters()
    piv = c1["pivots"] - c0["pivots"]
    print(f"K={K:5d} ms/launch {[round(v, 3) for v in ms]}  pivots {piv}  Mpiv/s {piv / sum(ms) / 1e3:.1f}", flush=True)
    eng.close()
