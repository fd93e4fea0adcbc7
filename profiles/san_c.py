import sys, os
sys.path.insert(0, os.getcwd())
from dwsolver_b200 import problem as P
from dwsolver_b200.engine import Engine
K=int(sys.argv[1]); m=int(sys.argv[2]); n=int(sys.argv[3]); G=int(sys.argv[4]); it=int(sys.argv[5])
prob = P.gen_dense(K, m, n, 16, 64001, dens_a=0.02)
eng = Engine(prob, force_path=4, cluster_size=G, max_iter_mult=it, no_dual_start=int(os.environ.get('NO_DUAL', '0')))
pr = eng.initial()
print("status", [p["status"] for p in pr], eng.counters())
