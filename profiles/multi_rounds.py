#!/usr/bin/env python
"""DW rounds / master iterations / seconds of dw_solve_blocks with DWSOLVER_MULTI = 1, 2, 3, 4 dual vectors per round
(SURVEY 8f-4).  usage: python profiles/multi_rounds.py [A|B]"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dwsolver_b200 import problem as P
from dwsolver_b200.solver import dw_solve_blocks
tag = sys.argv[1] if len(sys.argv) > 1 else "B"
prob = P.make_config(tag)
os.chdir(tempfile.mkdtemp())
dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)        # warm the process
for multi in (1, 2, 3, 4):
    os.environ["DWSOLVER_MULTI"] = str(multi)
    r = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    st = r["stats"]
    print(json.dumps(dict(config=tag, dual_vectors_per_round=multi, dw_rounds=st["dw_iterations"], master_columns=st["master_columns"],
                          seconds=round(st["seconds_total"], 4), subproblem_seconds=round(st["seconds_subproblems"], 4),
                          master_seconds=round(st["seconds_master"], 4), sub_pivots=st["sub_pivots"], objective=r["objective"], status=r["status"])), flush=True)
