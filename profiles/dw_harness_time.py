#!/usr/bin/env python
"""DW time-to-optimal of a synthetic configuration with the TEST HARNESS master (scipy/HiGHS, tests/dw_loop.py)
and either the CUDA engine or the CPU oracle ("port": pthreads + oracle/, not GLPK) for the subproblems.
Used for config B, whose 8448-row restricted master is beyond the product's own dense-inverse master LP
(DESIGN.md §7-2).  Prints one JSON line: wall time split into subproblem rounds and master solves."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dwsolver_b200 import problem as P  # noqa: E402
from dw_loop import run_dw  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "B"
arm = sys.argv[2] if len(sys.argv) > 2 else "gpu"
prob = P.make_config(tag)
t_sub = [0.0]


class Timed:
    def __init__(self, inner):
        self.inner = inner

    def initial(self):
        t = time.perf_counter(); r = self.inner.initial(); t_sub[0] += time.perf_counter() - t; return r

    def iterate(self, pi, phase_one):
        t = time.perf_counter(); r = self.inner.iterate(pi, phase_one); t_sub[0] += time.perf_counter() - t; return r


if arm == "gpu":
    from dwsolver_b200.engine import Engine
    eng = Engine(prob)
    eng.initial(); eng.reset()          # warm the CUDA context, then start again from the all-logical basis
    inner = eng
    note = "CUDA engine (dwgpu_iterate, dense host mailboxes incl. x_k)"
else:
    from oracle import oracle as O
    blocks = O.OracleBlocks(prob)
    threads = os.cpu_count()

    class Orc:
        def initial(self):
            return blocks.initial(threads)

        def iterate(self, pi, phase_one):
            return blocks.iterate(pi, phase_one, threads)
    inner = Orc()
    note = f"CPU port: {threads} pthreads + oracle/ (NOT GLPK)"
t0 = time.perf_counter()
res = run_dw(prob, Timed(inner))
wall = time.perf_counter() - t0
print(json.dumps(dict(config=tag, arm=arm, subproblems=note, master="scipy/HiGHS (test harness)", status=res["status"],
                      objective=res["objective"], dw_iterations=res["iterations"], wall_s=round(wall, 3),
                      subproblem_s=round(t_sub[0], 3), master_and_python_s=round(wall - t_sub[0], 3))))
