"""The host side of the drop-in on a machine WITHOUT a GPU: dw_solve's Phase I / Phase II loop, stop rules, block
sharding over several engines, the container path and the command line (dwsolver_b200/host), run against a TEST
DOUBLE of the engine ABI (tests/stub_engine/stub_dwgpu.c: the ten entry points the host library imports, on the CPU
oracle).  The double lives in tests/_build/ next to COPIES of the host library and of the CLI — their $ORIGIN
run-path finds it there; the product's own lib/ directory only ever holds the CUDA engine, which has no CPU
fallback (tests/test_abi.py).  What these tests pin is the host logic; the subproblem numerics are the oracle's, and
the GPU engine is compared with that oracle in tests/test_gpu_parity.py (-m gpu), where the same CLI checks run
against the real engine (tests/test_host_library.py).

Everything goes through the CLI in a child process: a second libdwgpu.so cannot be loaded into a Python process that
may already hold the real one.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import EXAMPLES, ROOT, load_example
from dwsolver_b200 import problem as P

BUILD = os.path.join(ROOT, "tests", "_build")
CLI = os.path.join(BUILD, "bin", "dwsolver")


@pytest.fixture(scope="module", autouse=True)
def stub_build():
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "stub_engine")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    ldd = subprocess.run(["ldd", CLI], capture_output=True, text=True).stdout
    assert os.path.join("tests", "_build") in ldd.split("libdwgpu.so")[1].splitlines()[0]     # the double, not the engine


def run(cwd, *args, env=None):
    e = dict(os.environ, DWSOLVER_TEST_DOUBLE="tests-only")      # the double refuses to run unannounced
    e.pop("DWSOLVER_DEVICES", None)
    e.update(env or {})
    return subprocess.run([CLI, *args], cwd=str(cwd), capture_output=True, text=True, timeout=600, env=e)


def banner(out):
    lines = [ln for ln in out.stdout.splitlines() if "Master objective value" in ln]
    assert lines, out.stdout + out.stderr
    return lines[-1].split("=")[1].strip()


def solution(cwd):
    return dict(ln.split("\t") for ln in open(os.path.join(str(cwd), "relaxed_solution")).read().splitlines())


@pytest.mark.parametrize("name", EXAMPLES)
def test_examples_through_the_cli(name, golden, tmp_path):
    """tests/dw-tests.sh of the reference: objective banner (and its %e literal), sorted relaxed_solution (%f)"""
    P.write_block_angular(load_example(name), str(tmp_path))
    out = run(tmp_path, "-g", "guidefile")
    assert out.returncode == 0, out.stderr
    want = golden["objective"][name]
    assert abs(float(banner(out)) - want) <= 1e-6 * (1 + abs(want))
    if name in golden["objective_literal"]:
        assert banner(out) == golden["objective_literal"][name]
    sol = golden["relaxed_solution"].get(name)
    if sol:
        rs = solution(tmp_path)
        for nm, v in sol.items():
            assert "%f" % (float(rs[nm]) + 0.0) == "%f" % v, (nm, rs[nm], v)


@pytest.mark.parametrize("devices", ["0,0,0", "all", "0,0,0,0,0,0,0"])
def test_sharded_solve_equals_single_engine(devices, tmp_path):
    """SURVEY 8e inside dw_solve: contiguous shards, one host thread per shard and round — same banner, same
    relaxed_solution as one engine (per-block results do not depend on the shard a block sits in)"""
    prob = P.gen_mcf(40, 16, 40, 8, seed=3)
    P.write_block_angular(prob, str(tmp_path))
    one = run(tmp_path, "-g", "guidefile")
    sol_one = solution(tmp_path)
    os.remove(tmp_path / "relaxed_solution")
    many = run(tmp_path, "-g", "guidefile", "--devices", devices)
    assert one.returncode == 0 and many.returncode == 0, one.stderr + many.stderr
    assert banner(one) == banner(many)
    assert solution(tmp_path) == sol_one and len(sol_one) == sum(b.n for b in prob.blocks)
    # against the monolithic LP
    from highs_util import solve_monolithic
    ref = solve_monolithic(prob)["objective"]
    assert abs(float(banner(many)) - ref) <= 1e-6 * (1 + abs(ref))


def test_more_devices_than_blocks(golden, tmp_path):
    P.write_block_angular(load_example("book_dantzig"), str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", env={"DWSOLVER_DEVICES": "0,0,0,0,0,0,0,0"})
    assert out.returncode == 0 and banner(out) == golden["objective_literal"]["book_dantzig"]


@pytest.mark.parametrize("name", ["four_sea", "book_lasdon"])
def test_container_solve_equals_text_solve(name, tmp_path):
    P.write_block_angular(load_example(name), str(tmp_path))
    text = run(tmp_path, "-g", "guidefile")
    sol_text = solution(tmp_path)
    os.remove(tmp_path / "relaxed_solution")
    assert run(tmp_path, "-g", "guidefile", "--write-container", "p.dwb", "--quiet").returncode == 0
    box = run(tmp_path, "--container", "p.dwb")
    assert text.returncode == 0 and box.returncode == 0, text.stderr + box.stderr
    assert banner(text) == banner(box) and solution(tmp_path) == sol_text      # original variable names travel
    assert run(tmp_path, "--container", "missing.dwb", "--quiet").returncode == 2   # DW_STATUS_ERR_FILE


def test_engine_failures_become_status_99(tmp_path):
    """the CUDA layer never exits the process: a failing dwgpu_create / dwgpu_iterate_packed comes back as
    DW_STATUS_ERR_INTERNAL with the engine's message on stderr (error conventions)"""
    P.write_block_angular(load_example("web_trick"), str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", "--quiet", env={"DWSTUB_FAIL_CREATE": "1"})
    assert out.returncode == 99 and "creation failure requested" in out.stderr
    out = run(tmp_path, "-g", "guidefile", "--quiet", env={"DWSTUB_FAIL_ITERATE": "1"})
    assert out.returncode == 99 and "round failure requested" in out.stderr
    out = run(tmp_path, "-g", "guidefile", "--quiet", "--devices", "0,0", env={"DWSTUB_FAIL_ITERATE": "1"})
    assert out.returncode == 99


def test_a_block_that_did_not_reach_an_optimum_never_becomes_a_column(tmp_path):
    """the kernels have real failure exits (status 1: iteration limit / numerical failure; 4: infeasible).
    A block reporting one mid-run — with an attractive objective, so that the acceptance test r_k - obj_k > 1e-6 would
    take it — must not enter the master: status 1 ends the solve with DW_STATUS_ERR_INTERNAL, status 4 is skipped."""
    P.write_block_angular(load_example("web_trick"), str(tmp_path))
    ref = run(tmp_path, "-g", "guidefile")
    assert ref.returncode == 0
    for rd in (1, 2):
        out = run(tmp_path, "-g", "guidefile", "--quiet", env={"DWSTUB_STATUS": f"{rd}:0:1"})
        assert out.returncode == 99 and "did not reach an optimum" in out.stderr, (rd, out.returncode, out.stderr)
    out = run(tmp_path, "-g", "guidefile", env={"DWSTUB_STATUS": "2:0:4"})
    assert out.returncode == 0 and "Subproblem 0 is infeasible." in out.stdout
    assert float(banner(out)) >= float(banner(ref)) - 1e-9      # built from real proposals only: never below the optimum


@pytest.mark.parametrize("name", ["book_dantzig", "four_sea", "web_trick", "book_bertsimas"])
@pytest.mark.parametrize("P_", [2, 4])
def test_several_dual_vectors_per_round_reach_the_same_optimum(name, P_, golden, tmp_path):
    """with DWSOLVER_MULTI=P every Phase II round prices the blocks with P dual vectors (the master's duals
    and points between them and the previous round's) through dwgpu_iterate_multi and offers all proposals to the
    master, accepted by their reduced cost at the master's duals.  Same optimum as the one-vector loop."""
    P.write_block_angular(load_example(name), str(tmp_path))
    ref = run(tmp_path, "-g", "guidefile")
    out = run(tmp_path, "-g", "guidefile", env={"DWSOLVER_MULTI": str(P_)})
    assert ref.returncode == 0 and out.returncode == 0, out.stderr
    want = golden["objective"][name]
    assert abs(float(banner(out)) - want) <= 1e-6 * (1 + abs(want))
    assert abs(float(banner(out)) - float(banner(ref))) <= 1e-9 * (1 + abs(want))


def test_the_double_refuses_to_run_unannounced(tmp_path):
    P.write_block_angular(load_example("single_sub"), str(tmp_path))
    env = {k: v for k, v in os.environ.items() if k != "DWSOLVER_TEST_DOUBLE"}
    out = subprocess.run([CLI, "-g", "guidefile", "--quiet"], cwd=str(tmp_path), capture_output=True, text=True, env=env)
    assert out.returncode == 99 and "test double" in out.stderr


def test_iteration_limits_and_missing_files(tmp_path):
    prob = P.gen_mcf(24, 12, 30, 6, seed=11)
    P.write_block_angular(prob, str(tmp_path))
    full = run(tmp_path, "-g", "guidefile", "--quiet")
    assert full.returncode == 0
    assert run(tmp_path, "-g", "guidefile", "--quiet", "--phase2_max", "1").returncode == 5    # DW_STATUS_ERR_PHASE2_LIMIT
    assert run(tmp_path, "-g", "guidefile", "--quiet", "--phase1_max", "1").returncode == 4    # DW_STATUS_ERR_PHASE1_LIMIT
    os.remove(tmp_path / "sub3.lp")
    assert run(tmp_path, "-g", "guidefile", "--quiet").returncode == 2                         # DW_STATUS_ERR_FILE
    assert run(tmp_path, "-g", "no_such_guidefile").returncode == 1


def _tiny_block(lb, ub, c):
    return P.Block(m=1, n=1, a_rowptr=np.array([0, 1], np.int32), a_col=np.array([0], np.int32), a_val=np.array([1.0]),
                   row_lb=np.array([-np.inf]), row_ub=np.array([100.0]), col_lb=np.array([lb]), col_ub=np.array([ub]),
                   c_file=np.array([c]), c_master=np.array([c]), d_row=np.array([0], np.int32),
                   d_col=np.zeros(1, np.int32), d_val=np.array([1.0]))


@pytest.mark.parametrize("case,want_rc", [("infeasible", 3), ("feasible", 0), ("unbounded", 6)])
def test_status_codes(case, want_rc, tmp_path):
    """src/dw_solver.h:88-97 — infeasible linking row -> DW_STATUS_ERR_INFEASIBLE with the reference's message
    (src/dw_solver.c:516-521); a block unbounded below -> DW_STATUS_ERR_UNBOUNDED (src/dw_phases.c:111-115)"""
    if case == "unbounded":
        blocks, ub = [_tiny_block(-np.inf, 5.0, 1.0)], 3.0              # min x, x free below
    else:
        blocks = [_tiny_block(1.0, 5.0, 1.0), _tiny_block(1.0, 5.0, 1.0)]   # x1, x2 >= 1 and x1 + x2 <= ub
        ub = 1.0 if case == "infeasible" else 3.0
    prob = P.BlockAngularLP(K=len(blocks), L=1, blocks=blocks, link_lb=np.array([-np.inf]), link_ub=np.array([ub]))
    P.write_block_angular(prob, str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", "--quiet" if case != "feasible" else "--output-normal")
    assert out.returncode == want_rc, out.stdout + out.stderr
    if case == "infeasible":
        assert "Problem is infeasible." in out.stdout
    if case == "unbounded":
        assert "unbounded" in out.stderr
    if case == "feasible":
        assert abs(float(banner(out)) - 2.0) <= 1e-9


_LIB_API_CHILD = r"""
import ctypes as C, math, os, sys
sys.path.insert(0, {root!r})
from dwsolver_b200.solver import DwOptions, DwResult
L = C.CDLL({lib!r})
L.dw_version.restype = C.c_char_p
L.dw_solve.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(DwOptions), C.POINTER(DwResult)]
master, subs = {master!r}.encode(), [s.encode() for s in {subs!r}]
arr = (C.c_char_p * len(subs))(*subs)
o = DwOptions(); L.dw_options_init(C.byref(o))
# T1 defaults (tests/test_lib_api.c:74-94)
assert (o.verbosity, o.max_phase1_iterations, o.max_phase2_iterations, o.rounding_flag, o.integerize_flag) == (5, 100, 3000, 0, 0)
assert abs(o.mip_gap - 0.01) < 1e-10 and o.shift == 0.0 and o.print_relaxed_sol == 1 and o.perturb == 0
# T2/T3 solve and free (:96-125)
o.verbosity = -1
res = DwResult()
rc = L.dw_solve(master, arr, len(subs), C.byref(o), C.byref(res))
assert rc == 0 and res.status == 0 and res.num_vars > 0 and bool(res.x) and math.isfinite(res.objective_value)
first = res.objective_value
L.dw_result_free(C.byref(res))
assert not res.x and res.num_vars == 0
# T4 NULL master (:127-139)
res = DwResult()
assert L.dw_solve(None, arr, len(subs), C.byref(o), C.byref(res)) == 1 and not res.x
# T5 version (:144-150)
assert L.dw_version().startswith(b"dwsolver")
# T6 silent verbosity prints nothing (:157-190), T7 custom mip_gap (:195-213)
o.verbosity = -1; o.mip_gap = 0.05
res = DwResult()
assert L.dw_solve(master, arr, len(subs), C.byref(o), C.byref(res)) == 0 and res.objective_value == first
L.dw_result_free(C.byref(res))
# T8 NULL options = defaults (:218-240); a fourth solve in one process reaches the same optimum
res = DwResult()
assert L.dw_solve(master, arr, len(subs), None, C.byref(res)) == 0 and res.objective_value == first
L.dw_result_free(C.byref(res))
print("LIBAPI-OK %.10g" % first)
"""


def test_library_api_like_the_reference(golden, tmp_path):
    """tests/test_lib_api.c of the reference (tests 1-8: defaults, solve + free, NULL master, version, silent
    verbosity, custom mip_gap, NULL options, four dw_solve calls in one process) on the host library, in a child
    process whose libdwgpu.so is the test double"""
    import sys
    paths = P.write_block_angular(load_example("book_bertsimas"), str(tmp_path))
    code = _LIB_API_CHILD.format(root=ROOT, lib=os.path.join(BUILD, "lib", "libdwsolver_b200.so"),
                                 master=paths["master"], subs=paths["subs"])
    out = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, DWSOLVER_TEST_DOUBLE="tests-only"))
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert lines[-1].startswith("LIBAPI-OK")
    want = golden["objective"]["book_bertsimas"]
    assert abs(float(lines[-1].split()[1]) - want) <= 1e-7 * (1 + abs(want))
    # T6: with verbosity -1 (silent) the only solve that printed is the NULL-options one (default verbosity 5)
    silent_part = out.stdout.split("####")[0]
    assert "Master objective" not in silent_part


REFERENCE_EXAMPLES = "/root/reference/examples"
# tests/dw-tests.sh of the reference, run for run: (example directory, guidefile, how it is judged)
DW_TESTS_SH = [("book_bertsimas", "guidefile", "solution"), ("book_lasdon", "guidefile", "solution"),
               ("web_mitchell", "guidefile", "solution"), ("web_trick", "guidefile", "-4.000000e+01"),
               ("four_sea", "guidefile", "1.200000e+01"), ("book_dantzig", "guidefile", "6.357895e+01"),
               ("single_sub", "guidefile", "solution"), ("one_iter", "guidefile", "-6.000000e+00"),
               ("neg_y", "guidefile", "solution")]


@pytest.mark.skipif(not os.path.isdir(REFERENCE_EXAMPLES), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("example,guide,judge", DW_TESTS_SH)
def test_reference_test_script_on_the_reference_files(example, guide, judge, tmp_path):
    """The reference's own acceptance script (tests/dw-tests.sh:11-160) on the reference's own input files — guidefile
    and CPLEX-LP text exactly as they lie under examples/ (copied to a scratch directory, the tree is read-only), through
    this build's CLI: `sort relaxed_solution | diff -w -B ex_relaxed_solution -` for the five solution-judged runs,
    the last `Master objective value` banner literal for the other four."""
    import shutil
    work = tmp_path / example
    shutil.copytree(os.path.join(REFERENCE_EXAMPLES, example), work)
    if judge == "solution":
        out = run(work, "--no-write-final-master", "--quiet", "-g", guide)
        assert out.returncode == 0, out.stdout + out.stderr
        got = sorted(open(work / "relaxed_solution").read().splitlines())
        norm = lambda lines: [" ".join(ln.split()) for ln in lines if ln.strip()]          # diff -w -B
        assert norm(got) == norm(open(work / "ex_relaxed_solution").read().splitlines())
    else:
        out = run(work, "-g", guide)
        assert out.returncode == 0, out.stdout + out.stderr
        assert judge in [ln for ln in out.stdout.splitlines() if "Master objective value" in ln][-1]


def test_random_instances_match_highs_through_the_host_loop():
    """profiles/stress_dw.py — 24 random block-angular instances (multicommodity flow of two sizes, dense blocks) through
    dw_solve_blocks against the HiGHS optimum of the monolithic LP, 1e-7 relative — with the host library bound to the
    engine test double (child process; the stress script itself is what `-m gpu` runs against the real engine)"""
    import sys
    code = ("import sys, runpy\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import dwsolver_b200.solver as S\n"
            f"S.LIB_PATH = {os.path.join(BUILD, 'lib', 'libdwsolver_b200.so')!r}\n"
            "sys.argv = ['stress_dw.py', '24']\n"
            f"runpy.run_path({os.path.join(ROOT, 'profiles', 'stress_dw.py')!r}, run_name='__main__')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, DWSOLVER_TEST_DOUBLE="tests-only"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "24 / 24 instances match HiGHS" in out.stdout


@pytest.mark.parametrize("name", ["book_dantzig", "four_sea", "web_mitchell"])
def test_final_master_file_solves_to_the_same_optimum(name, golden, tmp_path):
    """--write-final-master (src/dw_solver.c:749-753): done.cpxlp holds the final restricted master — the original
    linking-row names, sub<k>_convexity rows, every accepted lambda — and "solving this on its own later will provide
    the optimal value": read back with the package's LP reader and solved by HiGHS it gives the banner's objective"""
    import scipy.sparse as sp
    from scipy.optimize import linprog
    prob = load_example(name)
    P.write_block_angular(prob, str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", "--write-final-master")
    assert out.returncode == 0 and "Printing final master to done.cpxlp" in out.stdout
    lp = P.read_lp(str(tmp_path / "done.cpxlp"))
    assert lp.m == prob.L + prob.K and lp.row_names[prob.L:] == [f"sub{k + 1}_convexity" for k in range(prob.K)]
    assert sum(nm.startswith("lambda_") for nm in lp.col_names) >= prob.K
    A = sp.csr_matrix((lp.a_vals, (lp.a_rows, lp.a_cols)), shape=(lp.m, lp.n))
    assert np.array_equal(lp.row_lb, lp.row_ub)                                   # all rows are equalities by now
    res = linprog(lp.obj, A_eq=A, b_eq=lp.row_ub, bounds=list(zip(lp.col_lb, [None if not np.isfinite(u) else u for u in lp.col_ub])),
                  method="highs")
    assert res.status == 0
    want = golden["objective"][name]
    assert abs(res.fun + float(prob.shift or 0.0) - want) <= 1e-7 * (1 + abs(want))
    assert abs(float(banner(out)) - want) <= 1e-6 * (1 + abs(want))
    # without the flag nothing is written
    os.remove(tmp_path / "done.cpxlp")
    assert run(tmp_path, "-g", "guidefile", "--quiet").returncode == 0 and not os.path.exists(tmp_path / "done.cpxlp")


@pytest.mark.parametrize("name", EXAMPLES)
def test_perturbed_solve_reaches_the_same_optimum(name, golden, tmp_path):
    """-p / --perturb (src/dw_solver.c:427-431, 701-724): linking right-hand sides moved by 1e-8 * i during the solve,
    moved back and the master re-solved at the end — the final banner is the unperturbed optimum"""
    P.write_block_angular(load_example(name), str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", "--perturb")
    assert out.returncode == 0, out.stdout + out.stderr
    want = golden["objective"][name]
    assert abs(float(banner(out)) - want) <= 1e-6 * (1 + abs(want))


@pytest.mark.parametrize("block", ["default", "0"])
def test_wide_master_against_highs(block, tmp_path):
    """128 linking rows: the master's working inverse is updated in blocks of 8 by default (host/gub_master.cc);
    DWSOLVER_MASTER_BLOCK=0 is the unblocked solver — both reach the HiGHS optimum of the monolithic LP"""
    from highs_util import solve_monolithic
    prob = P.gen_mcf(48, 128, 300, 128, seed=21)
    P.write_block_angular(prob, str(tmp_path))
    out = run(tmp_path, "-g", "guidefile", env={} if block == "default" else {"DWSOLVER_MASTER_BLOCK": block})
    assert out.returncode == 0, out.stdout + out.stderr
    ref = solve_monolithic(prob)["objective"]
    assert abs(float(banner(out)) - ref) <= 1e-6 * (1 + abs(ref))
