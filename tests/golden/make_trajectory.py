#!/usr/bin/env python
"""Record the master-dual trajectory of a real Dantzig-Wolfe run on a synthetic configuration.

    python tests/golden/make_trajectory.py A            # -> tests/golden/traj_A.npz
    python tests/golden/make_trajectory.py B --gpu      # subproblems on the CUDA engine (GPU box)

The restricted master is solved with scipy/HiGHS (tests/dw_loop.py, restating SURVEY.md Appendix
A.2-A.5); the subproblems with the CPU oracle (default) or the CUDA engine.  The file holds, per
round r: the duals pi_r (L doubles) and the phase flag, exactly what the reference's master
publishes to its workers (src/dw_phases.c:271-275, 526-530), plus the objective the run reached.
bench.py replays this sequence so that every timed step does the work of one complete DW solve.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from dwsolver_b200 import problem as P  # noqa: E402
from dw_loop import run_dw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--monolithic", action="store_true", help="also solve the monolithic LP with HiGHS")
    args = ap.parse_args()
    t0 = time.time()
    prob = P.make_config(args.tag)
    print(f"generated {args.tag}: K={prob.K} L={prob.L} in {time.time()-t0:.1f}s", flush=True)
    if args.gpu:
        from dwsolver_b200.engine import Engine
        eng = Engine(prob)
        count = lambda: (lambda c: (c["pivots"], c["flips"]))(eng.counters())
    else:
        from oracle import oracle as O
        blocks = O.OracleBlocks(prob)

        class Eng:
            def initial(self):
                return blocks.initial(args.threads)

            def iterate(self, pi, phase_one):
                return blocks.iterate(pi, phase_one, args.threads)
        eng = Eng()
        count = lambda: blocks.counters()[:2]
    rec = []
    log = []

    class Logged:
        def initial(self):
            t = time.time(); r = eng.initial(); log.append(count()); print("initial", log[-1], f"{time.time()-t:.1f}s", flush=True); return r

        def iterate(self, pi, phase_one):
            t = time.time(); r = eng.iterate(pi, phase_one); log.append(count())
            print(f"round {len(log)-1} phase_one={phase_one} cum(piv,flip)={log[-1]} {time.time()-t:.1f}s", flush=True); return r
    res = run_dw(prob, Logged(), record=rec, verbose=True)
    print(res["status"], res["iterations"], res["objective"], f"{time.time()-t0:.1f}s")
    out = dict(pis=np.array([r[0] for r in rec]), phase_one=np.array([r[1] for r in rec], np.int32),
               objective=np.float64(res["objective"]), cum_counts=np.array(log, np.int64),
               status=np.array(res["status"]))
    if args.monolithic:
        from highs_util import solve_monolithic
        mono = solve_monolithic(prob)
        print("monolithic HiGHS objective", mono["objective"])
        out["monolithic_objective"] = np.float64(mono["objective"])
    np.savez_compressed(os.path.join(HERE, f"traj_{args.tag}.npz"), **out)


if __name__ == "__main__":
    main()
