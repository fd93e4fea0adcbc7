#!/usr/bin/env python
"""The product's restricted-master solver (host/gub_master.cc) at config scale WITHOUT a GPU: the DW loop of
tests/dw_loop.py with the CPU oracle for the subproblems and ProductMaster for the master.  Prints the master's
seconds and iteration count; DWSOLVER_MASTER_DEBUG=1 adds the solver's own per-phase breakdown per call."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dwsolver_b200 import problem as P  # noqa: E402
from dw_loop import ProductMaster, run_dw  # noqa: E402
from oracle import oracle as O  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "B"
prob = P.make_config(tag)
blocks = O.OracleBlocks(prob)
threads = os.cpu_count()


class Orc:
    def initial(self):
        return blocks.initial(threads)

    def iterate(self, pi, phase_one):
        return blocks.iterate(pi, phase_one, threads)


m = ProductMaster(prob)
t0 = time.perf_counter()
res = run_dw(prob, Orc(), master=m)
print(json.dumps(dict(config=tag, status=res["status"], objective=res["objective"], dw_iterations=res["iterations"],
                      wall_s=round(time.perf_counter() - t0, 2), master_s=round(m.seconds, 3), master_iterations=m.iterations)))
