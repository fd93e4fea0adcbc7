#!/usr/bin/env python
"""Config A (or any) replay timing under a forced kernel path: python profiles/scan_path.py A 1 3"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

tag = sys.argv[1]
pis, phases, _ = load_trajectory(tag)
pis_dev = torch.from_numpy(pis).cuda()
prob = P.make_config(tag)
for path in [int(a) for a in sys.argv[2:]]:
    eng = Engine(prob, fetch_x=False, force_path=path)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        eng.reset(); eng.sync()
        c0 = eng.counters()
        ev0.record()
        ms = []
        eng.initial_solve_device()
        for r in range(len(phases)):
            eng.iterate_device(pis_dev[r].data_ptr(), phases[r])
        eng.sync()
        c1 = eng.counters()
    # per-launch kernel times of the last replay need a sync per launch: separate pass
    eng.reset(); ms = []
    eng.initial_solve_device(); ms.append(max(eng.last_kernel_ms()))
    for r in range(len(phases)):
        eng.iterate_device(pis_dev[r].data_ptr(), phases[r]); ms.append(max(eng.last_kernel_ms()))
    piv = c1["pivots"] + c1["flips"] - c0["pivots"] - c0["flips"]
    print(f"path {path}: kernel ms/launch {[round(v, 3) for v in ms]} sum {sum(ms):.3f} ms, pivots+flips {piv}, {piv / sum(ms) / 1e3:.1f} Mpiv/s", flush=True)
    eng.close()
