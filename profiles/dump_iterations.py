#!/usr/bin/env python
"""Per-block simplex iterations of every round of a config's recorded DW trajectory -> gpurun_out/iters_<tag>.npy
(input of the list-scheduling study in profiles/lpt_study.py)."""
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
prob = P.make_config(tag)
eng = Engine(prob, fetch_x=False)
pis_dev = torch.from_numpy(pis).cuda()
eng.reset()
rows = []
eng.initial_solve_device(); eng.sync()
rows.append(eng.block_iterations().copy())
for r in range(len(phases)):
    eng.iterate_device(pis_dev[r].data_ptr(), phases[r]); eng.sync()
    rows.append(eng.block_iterations().copy())
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/iters_{tag}.npy", np.array(rows, dtype=np.int32))
print("saved", np.array(rows).shape)
eng.close()
