import os, sys, tempfile, json, time
sys.path.insert(0, "/root/repo")
from dwsolver_b200 import problem as P
from dwsolver_b200.solver import dw_solve_blocks
tag = sys.argv[1] if len(sys.argv) > 1 else "B"
prob = P.make_config(tag)
os.chdir(tempfile.mkdtemp())
for mk in (["ipm", "ipm", "gub"] if tag == "B" else ["ipm", "gub"]):
    os.environ["DWSOLVER_MASTER"] = mk
    t0 = time.perf_counter()
    r = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    st = r["stats"]
    print(mk, json.dumps(dict(wall=round(time.perf_counter() - t0, 3), seconds=st["seconds_total"], sub=st["seconds_subproblems"], master=st["seconds_master"],
                          rounds=st["dw_iterations"], status=r["status"], objective=r["objective"])), flush=True)
