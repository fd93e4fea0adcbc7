#!/usr/bin/env python
"""Config C (64 blocks x 2048 x 4096, cluster-per-block kernel): cold solve + a few priced rounds, timing,
counters and a HiGHS cross-check of some blocks.  usage: run_c.py [K] [rounds] [check_blocks]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.time()
prob = P.make_config("C", block_range=range(K))
print(f"generated C[{K}] in {time.time()-t0:.1f}s", flush=True)
eng = Engine(prob)
print("paths", set(eng.block_paths()), flush=True)
t0 = time.time()
props = eng.initial()
print(f"cold solve {time.time()-t0:.2f}s kernel {eng.last_kernel_ms()} ms", eng.counters(), flush=True)
print("status", sorted(set(p["status"] for p in props)), "iters/block", eng.block_iterations()[:8], flush=True)
if ncheck:
    from highs_util import solve_block
    for k in range(min(ncheck, K)):
        b = prob.blocks[k]
        ref = solve_block(b, b.c_file)
        print(f"block {k}: gpu obj {props[k]['obj']:.12e} highs {ref.fun:.12e} rel diff "
              f"{abs(props[k]['obj']-ref.fun)/(1+abs(ref.fun)):.2e}", flush=True)
rng = np.random.default_rng(1)
for r in range(rounds):
    pi = -rng.random(prob.L) * 0.05 * (r + 1)
    c0 = eng.counters()
    t0 = time.time()
    props = eng.iterate(pi, False)
    c1 = eng.counters()
    print(f"round {r}: {time.time()-t0:.2f}s kernel {eng.last_kernel_ms()} ms pivots {c1['pivots']-c0['pivots']} flips {c1['flips']-c0['flips']} "
          f"rebuilds {c1['rebuilds']-c0['rebuilds']} status {sorted(set(p['status'] for p in props))}", flush=True)
eng.close()
