#!/usr/bin/env python
"""Input path of dw_solve at config A (1024 block files + master, 12.9 MB of LP text), on the GPU box:
seconds_read of dw_solve with one reader thread and with the default team, and of dw_solve_container; the three
solves must reach the same objective.  Usage: python profiles/input_path_timing.py [A]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200 import solver as S  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
prob = P.make_config(tag)
d = tempfile.mkdtemp()
paths = P.write_block_angular(prob, d)
os.chdir(d)
out = {"config": tag, "host_cores": os.cpu_count(),
       "lp_text_bytes": sum(os.path.getsize(s) for s in paths["subs"]) + os.path.getsize(paths["master"])}
runs = []
for label, threads in (("warm-up", None), ("text, 1 reader", "1"), ("text, default readers", None)):
    if threads is None:
        os.environ.pop("DWSOLVER_READ_THREADS", None)
    else:
        os.environ["DWSOLVER_READ_THREADS"] = threads
    r = S.dw_solve(paths["master"], paths["subs"], verbosity=-1, print_relaxed_sol=0)
    runs.append(dict(run=label, status=r["status"], objective=r["objective"], seconds_read=round(r["stats"]["seconds_read"], 4),
                     seconds_total=round(r["stats"]["seconds_total"], 4)))
t0 = time.perf_counter()
rc = S.dw_write_container(paths["master"], paths["subs"], "instance.dwb")
out["write_container_s"] = round(time.perf_counter() - t0, 4)
out["container_bytes"] = os.path.getsize("instance.dwb")
r = S.dw_solve_container("instance.dwb", verbosity=-1, print_relaxed_sol=0)
runs.append(dict(run="container", status=r["status"], objective=r["objective"], seconds_read=round(r["stats"]["seconds_read"], 4),
                 seconds_total=round(r["stats"]["seconds_total"], 4)))
out["runs"] = runs
out["same_objective"] = len({x["objective"] for x in runs}) == 1 and rc == 0
print(json.dumps(out))
