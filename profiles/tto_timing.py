#!/usr/bin/env python
"""dw_solve_blocks on a config, three times in one process, with the driver's own section timers
(DWSOLVER_DEBUG_TIMING=1): where the time outside the master and the subproblem calls goes."""
import json
import os
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ["DWSOLVER_DEBUG_TIMING"] = "1"
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.solver import dw_solve_blocks  # noqa: E402

prob = P.make_config(sys.argv[1] if len(sys.argv) > 1 else "A")
os.chdir(tempfile.mkdtemp())
for _ in range(3):
    r = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    print(json.dumps(dict(status=r["status"], objective=r["objective"], **r["stats"])))
