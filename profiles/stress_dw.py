#!/usr/bin/env python
"""Stress: random block-angular instances through dw_solve_blocks (GPU subproblems + GUB master) against the HiGHS
optimum of the monolithic LP.  Prints one line per instance and a summary; exit code 1 on any mismatch."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.solver import dw_solve_blocks  # noqa: E402
from highs_util import solve_monolithic  # noqa: E402

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(2026)
bad = 0
os.chdir(tempfile.mkdtemp())
for t in range(n_inst):
    kind = t % 3
    if kind == 0:
        nodes = int(rng.choice([16, 32, 64])); K = int(rng.integers(8, 200)); cols = nodes * 2 + int(rng.integers(0, 20))
        L = int(rng.integers(2, nodes // 2))
        prob = P.gen_mcf(K, nodes, cols, L, seed=int(rng.integers(1, 10**6)))
        desc = f"mcf K={K} nodes={nodes} cols={cols} L={L}"
    elif kind == 1:
        nodes = int(rng.choice([128, 256])); K = int(rng.integers(8, 64)); cols = nodes * 2; L = int(rng.integers(8, 64))
        prob = P.gen_mcf(K, nodes, cols, L, seed=int(rng.integers(1, 10**6)))
        desc = f"mcf K={K} nodes={nodes} cols={cols} L={L}"
    else:
        K = int(rng.integers(2, 10)); m = int(rng.integers(20, 120)); n = 2 * m + int(rng.integers(0, 40)); L = int(rng.integers(2, 12))
        prob = P.gen_dense(K, m, n, L, seed=int(rng.integers(1, 10**6)), dens_a=0.08)
        desc = f"dense K={K} m={m} n={n} L={L}"
    t0 = time.time()
    out = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    dt = time.time() - t0
    ref = solve_monolithic(prob)["objective"]
    ok = out["status"] == 0 and ref is not None and abs(out["objective"] - ref) <= 1e-7 * (1 + abs(ref))
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} {desc}: status {out['status']} obj {out['objective']:.10g} highs {ref:.10g} "
          f"rounds {out['stats']['dw_iterations']} {dt:.2f}s", flush=True)
print(f"{n_inst - bad} / {n_inst} instances match HiGHS to 1e-7")
sys.exit(1 if bad else 0)
