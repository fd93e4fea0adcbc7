import os, sys, tempfile
sys.path.insert(0, "/root/repo")
from dwsolver_b200 import problem as P
from dwsolver_b200.solver import dw_solve_blocks
prob = P.make_config("B")
os.chdir(tempfile.mkdtemp())
r = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
print(r["stats"], r["objective"])
