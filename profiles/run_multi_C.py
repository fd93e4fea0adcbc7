#!/usr/bin/env python
"""Profiling driver: 4 blocks of config C's shape (2048 x 4096, D_k 25 % dense, cluster path), cold solve, then ONE
dwgpu_iterate_multi call with 16 dual vectors — the FP64 tensor-core pricing GEMM (price_multi_dmma_kernel) runs once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dwsolver_b200 import problem as P
from dwsolver_b200.engine import Engine
prob = P.make_config("C", block_range=range(4))
eng = Engine(prob, fetch_x=False)
print("dense pricing blocks:", eng.dense_pricing_blocks(), "paths", set(eng.block_paths()))
eng.initial_solve_device(); eng.sync()
rng = np.random.default_rng(1)
Pi = -rng.random((16, prob.L)) * 0.002
recs, ids = eng.iterate_multi(Pi, False)
print("status", [int((r["status"] == 5).all()) for r in recs])
eng.close()
