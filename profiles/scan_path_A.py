#!/usr/bin/env python
"""Config A through the shared-memory kernel (path 1) and through the HBM/L2-tableau kernels (path 3, CTA widths and the
warp kernel): device time of one replay of the recorded trajectory.  usage: python profiles/scan_path_A.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bench import load_trajectory
from dwsolver_b200 import problem as P
from dwsolver_b200.engine import Engine
pis, phases, _ = load_trajectory("A")
prob = P.make_config("A")
pis_dev = torch.from_numpy(pis).cuda()
for path, thr in ((1, ""), (3, "128"), (3, "64"), (3, "32")):
    if thr: os.environ["DWGPU_TG_THREADS"] = thr
    eng = Engine(prob, fetch_x=False, force_path=path)
    best = 1e9
    for rep in range(6):
        eng.reset(); eng.sync()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.initial_solve_device()
        for r in range(len(phases)):
            eng.iterate_device(pis_dev[r].data_ptr(), phases[r])
        eng.sync(); dt = time.perf_counter() - t0
        best = min(best, dt)
    c = eng.counters()
    print(f"path {path} threads {thr or '-'}: {best*1e3:.3f} ms per replay, {(c['pivots']+c['flips'])/6/best/1e6:.1f} M pivots/s, paths {set(eng.block_paths())}")
    eng.close()
