#!/usr/bin/env python
"""DW time-to-optimal through the drop-in public API (dw_solve, include/dwsolver.h) on a synthetic
configuration, beside the CPU arm (pthreads + oracle subproblems, HiGHS master: tests/dw_loop.py).

    python profiles/dw_time_to_optimal.py A [--cpu]

Writes the instance in the reference's input format (guidefile + CPLEX-LP files) to a temp directory,
calls dw_solve() once cold and prints one JSON line with the library's own timers (dw_last_stats)."""
import ctypes as C
import json
import os
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dwsolver_b200 import problem as P  # noqa: E402


class DwOptions(C.Structure):
    _fields_ = [("verbosity", C.c_int), ("mip_gap", C.c_double), ("max_phase1_iterations", C.c_int),
                ("max_phase2_iterations", C.c_int), ("rounding_flag", C.c_int), ("integerize_flag", C.c_int),
                ("enforce_sub_integrality", C.c_int), ("print_timing_data", C.c_int), ("print_final_master", C.c_int),
                ("print_relaxed_sol", C.c_int), ("perturb", C.c_int), ("shift", C.c_double)]


class DwResult(C.Structure):
    _fields_ = [("status", C.c_int), ("objective_value", C.c_double), ("num_vars", C.c_int), ("x", C.POINTER(C.c_double))]


class DwStats(C.Structure):
    _fields_ = [("dw_iterations", C.c_int), ("phase1_iterations", C.c_int), ("master_columns", C.c_int),
                ("sub_pivots", C.c_longlong), ("seconds_total", C.c_double), ("seconds_read", C.c_double),
                ("seconds_subproblems", C.c_double), ("seconds_master", C.c_double)]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "A"
    K = int(os.environ.get("DW_K", "0")) or None
    prob = P.make_config(tag, K=K)
    out = {"config": tag, "K": prob.K, "L": prob.L}
    if "--no-files" not in sys.argv:        # the reference's input path: guidefile + CPLEX-LP text
        with tempfile.TemporaryDirectory() as d:
            t0 = time.time()
            paths = P.write_block_angular(prob, d)
            out["write_lp_files_s"] = round(time.time() - t0, 2)
            L = C.CDLL(os.path.join(ROOT, "dwsolver_b200", "lib", "libdwsolver_b200.so"))
            L.dw_solve.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(DwOptions), C.POINTER(DwResult)]
            cwd = os.getcwd()
            os.chdir(d)
            try:
                for rep in range(2):                       # the second solve has a warm CUDA context
                    o = DwOptions()
                    L.dw_options_init(C.byref(o))
                    o.verbosity = -1
                    o.print_relaxed_sol = 0
                    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
                    res = DwResult()
                    t0 = time.time()
                    rc = L.dw_solve(paths["master"].encode(), subs, prob.K, C.byref(o), C.byref(res))
                    wall = time.time() - t0
                    st = DwStats()
                    L.dw_last_stats(C.byref(st))
                    out[f"gpu_solve_{rep}"] = dict(status=rc, objective=res.objective_value, wall_s=round(wall, 4),
                                                   read_s=round(st.seconds_read, 4), subproblems_s=round(st.seconds_subproblems, 4),
                                                   master_s=round(st.seconds_master, 4), dw_iterations=st.dw_iterations,
                                                   master_columns=st.master_columns, sub_pivots=st.sub_pivots)
                    L.dw_result_free(C.byref(res))
            finally:
                os.chdir(cwd)
    if "--arrays" in sys.argv:                  # in-memory entry point: no LP text at all
        from dwsolver_b200.solver import dw_solve_blocks
        for rep in range(2):
            t0 = time.time()
            r = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
            st = r["stats"]
            out[f"gpu_blocks_{rep}"] = dict(status=r["status"], objective=r["objective"], wall_s=round(time.time() - t0, 4),
                                            solve_s=round(st["seconds_total"], 4), subproblems_s=round(st["seconds_subproblems"], 4),
                                            master_s=round(st["seconds_master"], 4), dw_iterations=st["dw_iterations"],
                                            master_columns=st["master_columns"], sub_pivots=st["sub_pivots"])
    if "--cpu" in sys.argv:
        from dw_loop import run_dw
        from oracle import oracle as O
        threads = os.cpu_count()
        blocks = O.OracleBlocks(prob)

        class Eng:
            def initial(self):
                return blocks.initial(threads)

            def iterate(self, pi, phase_one):
                return blocks.iterate(pi, phase_one, threads)
        t0 = time.time()
        res = run_dw(prob, Eng())
        out["cpu_port"] = dict(objective=res["objective"], wall_s=round(time.time() - t0, 3), iterations=res["iterations"],
                               threads=threads, note="pthreads + CPU restatement subproblems, HiGHS master (tests/dw_loop.py); NOT GLPK")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
