"""Host side of the drop-in (dwsolver_b200/host -> libdwsolver_b200.so, bin/dwsolver).

CPU part: the C++ CPLEX-LP reader against the package's reader on LP text written from every golden
example, the master LP solver against HiGHS, the public symbols and struct layouts of
include/dwsolver.h (the reference's src/dw_solver.h:88-183).
GPU part (-m gpu): the reference's own test strategy (tests/dw-tests.sh) — run the CLI on a guidefile
and compare the `Master objective value` banner / the sorted relaxed_solution with the goldens; and
tests/test_lib_api.c — dw_solve() through the C API, bad-argument and missing-file status codes.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from conftest import EXAMPLES, ROOT, load_example
from dwsolver_b200 import problem as P

LIB = os.path.join(ROOT, "dwsolver_b200", "lib", "libdwsolver_b200.so")
HOOKS = os.path.join(ROOT, "dwsolver_b200", "lib", "libdwhost_testhooks.so")      # dwhost_* test entry points (tests/host_hooks)
CLI = os.path.join(ROOT, "dwsolver_b200", "bin", "dwsolver")
c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


class DwOptions(C.Structure):          # include/dwsolver.h, same order as src/dw_solver.h:105-118
    _fields_ = [("verbosity", C.c_int), ("mip_gap", C.c_double), ("max_phase1_iterations", C.c_int),
                ("max_phase2_iterations", C.c_int), ("rounding_flag", C.c_int), ("integerize_flag", C.c_int),
                ("enforce_sub_integrality", C.c_int), ("print_timing_data", C.c_int), ("print_final_master", C.c_int),
                ("print_relaxed_sol", C.c_int), ("perturb", C.c_int), ("shift", C.c_double)]


class DwResult(C.Structure):
    _fields_ = [("status", C.c_int), ("objective_value", C.c_double), ("num_vars", C.c_int), ("x", c_dp)]


def lib():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} is missing: run __graft_entry__.build()")
    L = C.CDLL(LIB)
    L.dw_version.restype = C.c_char_p
    L.dw_solve.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(DwOptions), C.POINTER(DwResult)]
    return L


def test_lp_reader_keywords_only_at_line_start(tmp_path):
    """glp_read_lp recognises section keywords only at the start of a line; columns named like keywords
    (st, end, bin, int, general, bounds) inside an objective or constraint line are columns."""
    path = tmp_path / "kw.lp"
    path.write_text("Minimize\n obj: 2 st + 3 end + bin - int + 5 general\nSubject To\n c1: st + end + bin >= 1\n"
                    " c2: int + general - bounds <= 4\nBounds\n 0 <= st <= 7\n end <= 3\n bounds free\nEnd\n")
    dims = (C.c_int * 3)()
    assert hooks().dwhost_lp_dims(str(path).encode(), dims) == 0
    assert list(dims) == [2, 6, 6]                       # 2 rows, 6 columns, 6 non-zeros


_HOOKS = None


def hooks():
    """the test-only entry points (tests/host_hooks/test_api.cc), a library of their own next to the product's"""
    global _HOOKS
    if _HOOKS is None:
        if not os.path.exists(HOOKS):
            pytest.fail(f"{HOOKS} is missing: run __graft_entry__.build()")
        _HOOKS = C.CDLL(HOOKS)
    return _HOOKS


def write_example(prob, d):
    """LP text + guidefile of a BlockAngularLP (the reference's input format) into directory d."""
    paths = P.write_block_angular(prob, str(d))
    return paths


def test_public_api_symbols_and_defaults():
    L = lib()
    import re
    header = open(os.path.join(ROOT, "include", "dwsolver.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dw_[a-z_]+)\s*\(", header))
    assert {"dw_options_init", "dw_solve", "dw_result_free", "dw_version", "dw_last_stats", "dw_solve_blocks",
            "dw_write_container", "dw_solve_container"} <= declared
    for name in declared:                                # every function include/dwsolver.h declares is exported
        assert hasattr(L, name), name
    o = DwOptions()
    L.dw_options_init(C.byref(o))
    assert (o.verbosity, o.max_phase1_iterations, o.max_phase2_iterations, o.print_relaxed_sol) == (5, 100, 3000, 1)
    assert (o.mip_gap, o.shift, o.print_final_master, o.perturb) == (0.01, 0.0, 0, 0)
    assert C.sizeof(DwResult) == 32 and C.sizeof(DwOptions) == 64
    assert b"dwsolver" in L.dw_version()


def test_bad_arguments_and_missing_files():
    L = lib()
    res = DwResult()
    assert L.dw_solve(None, None, 0, None, C.byref(res)) == 1          # DW_STATUS_ERR_BAD_ARGS
    arr = (C.c_char_p * 1)(b"/nonexistent/sub.lp")
    o = DwOptions()
    L.dw_options_init(C.byref(o))
    o.verbosity = -1
    assert L.dw_solve(b"/nonexistent/master.lp", arr, 1, C.byref(o), C.byref(res)) == 2   # DW_STATUS_ERR_FILE
    assert res.status == 2


def test_block_files_read_by_several_threads(tmp_path):
    """the K block files are parsed by a team of host threads (the reference serialises them under
    glpk_mutex/master_mutex, src/dw_subprob.c:105-111, 203-258); the blocks handed to the engine are the same for
    any number of readers, and an unreadable block file is still DW_STATUS_ERR_FILE."""
    L = lib()
    prob = P.gen_mcf(24, 12, 30, 6, seed=11)
    paths = write_example(prob, tmp_path)
    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
    hooks().dwhost_read_problem_digest.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_double)]
    digests = []
    for readers in (1, 3, 8, 0):
        out = (C.c_double * 8)()
        assert hooks().dwhost_read_problem_digest(paths["master"].encode(), subs, prob.K, readers, out) == 0
        digests.append(list(out))
    assert all(d == digests[0] for d in digests)                        # bit-identical, order-sensitive checksums
    d = digests[0]
    assert d[0] == prob.L
    assert d[1] == sum(b.m for b in prob.blocks) and d[2] == sum(b.n for b in prob.blocks)
    assert d[4] == sum(len(b.d_val) for b in prob.blocks)
    # a missing block file: every reader count reports it, and dw_solve maps it to DW_STATUS_ERR_FILE
    bad = [s.encode() for s in paths["subs"]]
    bad[prob.K // 2] = str(tmp_path / "no_such_block.lp").encode()
    bad_arr = (C.c_char_p * prob.K)(*bad)
    for readers in (1, 4):
        assert hooks().dwhost_read_problem_digest(paths["master"].encode(), bad_arr, prob.K, readers, (C.c_double * 8)()) == 1
    o = DwOptions()
    L.dw_options_init(C.byref(o))
    o.verbosity = -1
    res = DwResult()
    assert L.dw_solve(paths["master"].encode(), bad_arr, prob.K, C.byref(o), C.byref(res)) == 2
    assert res.status == 2


def _container_api(L):
    L.dw_write_container.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_double, C.c_char_p]
    L.dw_solve_container.argtypes = [C.c_char_p, C.POINTER(DwOptions), C.POINTER(DwResult)]
    hooks().dwhost_container_digest.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    hooks().dwhost_read_problem_digest.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_double)]


@pytest.mark.parametrize("source", ["mcf", "book_dantzig", "web_trick"])
def test_container_holds_what_the_lp_text_path_reads(source, tmp_path):
    """the binary container carries bit for bit what the LP-text path (glp_read_lp + the by-name column
    mapping, src/dw_subprob.c:105-111, 203-258) hands to the engine; damaged files are DW_STATUS_ERR_FILE."""
    L = lib()
    _container_api(L)
    if source == "mcf":
        prob = P.gen_mcf(24, 12, 30, 6, seed=5)
    else:
        prob = load_example(source)
    paths = write_example(prob, tmp_path)
    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
    box = str(tmp_path / "instance.dwb").encode()
    assert L.dw_write_container(paths["master"].encode(), subs, prob.K, 2.5, box) == 0
    text, packed = (C.c_double * 8)(), (C.c_double * 8)()
    assert hooks().dwhost_read_problem_digest(paths["master"].encode(), subs, prob.K, 1, text) == 0
    k_out, name = C.c_int(0), C.create_string_buffer(256)
    assert hooks().dwhost_container_digest(box, packed, C.byref(k_out), name, 256) == 0
    assert list(text) == list(packed) and k_out.value == prob.K
    first = name.value.decode()
    assert first and first in open(paths["subs"][0]).read().split()           # the block's own column names travel
    # damaged containers: truncated at several places, wrong magic, missing file
    blob = open(box, "rb").read()
    for cut in (7, 31, len(blob) // 3, len(blob) - 1):
        bad = tmp_path / f"cut{cut}.dwb"
        bad.write_bytes(blob[:cut])
        assert hooks().dwhost_container_digest(str(bad).encode(), packed, None, None, 0) == 1
    (tmp_path / "magic.dwb").write_bytes(b"NOTADWB!" + blob[8:])
    assert hooks().dwhost_container_digest(str(tmp_path / "magic.dwb").encode(), packed, None, None, 0) == 1
    o = DwOptions()
    L.dw_options_init(C.byref(o))
    o.verbosity = -1
    res = DwResult()
    assert L.dw_solve_container(str(tmp_path / "nope.dwb").encode(), C.byref(o), C.byref(res)) == 2 and res.status == 2
    assert L.dw_solve_container(str(tmp_path / "cut31.dwb").encode(), C.byref(o), C.byref(res)) == 2
    assert L.dw_solve_container(None, C.byref(o), C.byref(res)) == 1
    assert L.dw_write_container(b"/nonexistent/master.lp", subs, prob.K, 0.0, box) == 2


def test_cli_writes_a_container(tmp_path):
    """`dwsolver -g guidefile --write-container f` converts without solving (no GPU involved)"""
    L = lib()
    _container_api(L)
    prob = P.gen_mcf(12, 10, 24, 5, seed=9)
    paths = write_example(prob, tmp_path)
    out = subprocess.run([CLI, "-g", paths["guidefile"], "--write-container", "p.dwb", "--quiet"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    packed, text = (C.c_double * 8)(), (C.c_double * 8)()
    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
    assert hooks().dwhost_container_digest(str(tmp_path / "p.dwb").encode(), packed, None, None, 0) == 0
    assert hooks().dwhost_read_problem_digest(paths["master"].encode(), subs, prob.K, 2, text) == 0
    assert list(packed) == list(text)


@pytest.mark.parametrize("name", EXAMPLES)
def test_cpp_lp_reader_matches_package_reader(name, tmp_path):
    L = lib()
    prob = load_example(name)
    paths = write_example(prob, tmp_path)
    for path in [paths["master"]] + paths["subs"]:
        ref = P.read_lp(path)
        dims = (C.c_int * 3)()
        assert hooks().dwhost_lp_dims(path.encode(), dims) == 0
        m, n, nnz = dims[0], dims[1], dims[2]
        assert (m, n, nnz) == (ref.m, ref.n, len(ref.a_vals))
        obj, lb, ub = np.zeros(n), np.zeros(n), np.zeros(n)
        rlb, rub = np.zeros(m), np.zeros(m)
        ar, ac, av = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        names = C.create_string_buffer(64 * (n + 1))
        mx = C.c_int(0)
        d = lambda a: a.ctypes.data_as(c_dp)
        i = lambda a: a.ctypes.data_as(c_ip)
        assert hooks().dwhost_lp_read(path.encode(), d(obj), d(lb), d(ub), d(rlb), d(rub), i(ar), i(ac), d(av), names,
                                len(names), C.byref(mx)) == 0
        assert names.value.decode().split("\n")[:-1] == list(ref.col_names)
        assert np.array_equal(obj, ref.obj) and np.array_equal(lb, ref.col_lb) and np.array_equal(ub, ref.col_ub)
        assert np.array_equal(rlb, ref.row_lb) and np.array_equal(rub, ref.row_ub)
        assert np.array_equal(ar, ref.a_rows) and np.array_equal(ac, ref.a_cols) and np.array_equal(av, ref.a_vals)


@pytest.mark.parametrize("seed", range(6))
def test_master_lp_matches_highs(seed):
    """random equality-form LPs with boxed and unboxed columns, solved cold and with a warm second stage"""
    L = lib()
    rng = np.random.default_rng(seed)
    m, n = 12 + 3 * seed, 40 + 10 * seed
    A = sp.random(m, n, density=0.3, random_state=seed, data_rvs=lambda k: rng.integers(-4, 5, k).astype(float)).tocsc()
    A.eliminate_zeros()
    x0 = rng.random(n)
    ub = np.where(rng.random(n) < 0.5, 1.0, 1e300)
    rhs = A @ x0
    cost = rng.integers(-5, 6, n).astype(float)
    cost[ub >= 1e300] = np.abs(cost[ub >= 1e300]) + 1.0          # keep it bounded
    ref = linprog(cost, A_eq=A, b_eq=rhs, bounds=[(0, None if u >= 1e300 else u) for u in ub], method="highs")
    assert ref.status == 0
    for n_first in (0, n // 2):
        x, y = np.zeros(n), np.zeros(m)
        obj, its = C.c_double(0), C.c_longlong(0)
        cp = A.indptr.astype(np.int32)
        ix = A.indices.astype(np.int32)
        rc = hooks().dwhost_master_solve(m, n, cp.ctypes.data_as(c_ip), ix.ctypes.data_as(c_ip), A.data.ctypes.data_as(c_dp),
                                   cost.ctypes.data_as(c_dp), ub.ctypes.data_as(c_dp), rhs.ctypes.data_as(c_dp), n_first,
                                   x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp), C.byref(obj), C.byref(its))
        assert rc == 0
        assert abs(obj.value - ref.fun) <= 1e-8 * (1 + abs(ref.fun))
        assert np.abs(A @ x - rhs).max() <= 1e-8 and x.min() >= -1e-9 and (x - np.minimum(ub, 1e300)).max() <= 1e-9
        d = cost - A.T @ y                                        # dual feasibility of the returned duals
        at_lb, at_ub = x <= 1e-9, (ub < 1e300) & (x >= ub - 1e-9)
        assert d[at_lb & ~at_ub].min(initial=0.0) >= -1e-7 and d[at_ub & ~at_lb].max(initial=0.0) <= 1e-7


def run_cli(tmp_path, prob, *flags):
    write_example(prob, tmp_path)
    return subprocess.run([CLI, "-g", "guidefile", *flags], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXAMPLES)
def test_cli_reaches_reference_objective(name, golden, tmp_path):
    """tests/dw-tests.sh: `dwsolver -g guidefile`, last `Master objective value` banner (7 significant digits)"""
    prob = load_example(name)
    out = run_cli(tmp_path, prob)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if "Master objective value" in ln]
    assert lines
    want = golden["objective"][name]
    got = float(lines[-1].split("=")[1])
    assert abs(got - want) <= 1e-6 * (1 + abs(want))
    if name in golden["objective_literal"]:
        assert lines[-1].split("=")[1].strip() == golden["objective_literal"][name]
    sol = golden["relaxed_solution"].get(name)
    if sol:                                                       # the five deterministic fixtures
        rs = dict(ln.split("\t") for ln in open(tmp_path / "relaxed_solution").read().splitlines())
        for nm, v in sol.items():
            assert "%f" % (float(rs[nm]) + 0.0) == "%f" % v, (nm, rs[nm], v)


@pytest.mark.gpu
def test_dw_solve_api_on_synthetic_mcf(tmp_path):
    """tests/test_lib_api.c: dw_solve through the C API (twice in one process), objective vs the monolithic LP"""
    from highs_util import solve_monolithic
    L = lib()
    prob = P.gen_mcf(40, 16, 40, 8, seed=3)
    paths = write_example(prob, tmp_path)
    ref = solve_monolithic(prob)["objective"]
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for _ in range(2):
            o = DwOptions()
            L.dw_options_init(C.byref(o))
            o.verbosity = -1
            subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
            res = DwResult()
            rc = L.dw_solve(paths["master"].encode(), subs, prob.K, C.byref(o), C.byref(res))
            assert rc == 0 and res.status == 0
            assert abs(res.objective_value - ref) <= 1e-7 * (1 + abs(ref))
            assert res.num_vars > 0 and np.isfinite(res.x[0])
            L.dw_result_free(C.byref(res))
            assert not res.x
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_dw_solve_container_matches_dw_solve(tmp_path):
    """the same instance through LP text and through the binary container: same objective, same master columns, same
    relaxed_solution file (original variable names), and the stored objective constant is applied"""
    L = lib()
    _container_api(L)
    prob = P.gen_mcf(40, 16, 40, 8, seed=3)
    paths = write_example(prob, tmp_path)
    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
    box = str(tmp_path / "instance.dwb").encode()
    assert L.dw_write_container(paths["master"].encode(), subs, prob.K, 0.0, box) == 0
    shifted = str(tmp_path / "shifted.dwb").encode()
    assert L.dw_write_container(paths["master"].encode(), subs, prob.K, 100.0, shifted) == 0
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        o = DwOptions()
        L.dw_options_init(C.byref(o))
        o.verbosity = -1
        r_text, r_box, r_shift = DwResult(), DwResult(), DwResult()
        assert L.dw_solve(paths["master"].encode(), subs, prob.K, C.byref(o), C.byref(r_text)) == 0
        sol_text = sorted(open("relaxed_solution").read().splitlines())
        os.remove("relaxed_solution")
        assert L.dw_solve_container(box, C.byref(o), C.byref(r_box)) == 0
        sol_box = sorted(open("relaxed_solution").read().splitlines())
        assert r_box.objective_value == r_text.objective_value and r_box.num_vars == r_text.num_vars
        assert sol_box == sol_text and len(sol_box) == sum(b.n for b in prob.blocks)
        assert L.dw_solve_container(shifted, C.byref(o), C.byref(r_shift)) == 0
        assert abs(r_shift.objective_value - (r_text.objective_value + 100.0)) <= 1e-9 * (1 + abs(r_text.objective_value))
        for r in (r_text, r_box, r_shift):
            L.dw_result_free(C.byref(r))
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0,0", "all", "0,0,0,0,0,0,0"])
def test_dw_solve_sharded_over_several_engines(devices, tmp_path, monkeypatch):
    """SURVEY 8e inside the drop-in: DWSOLVER_DEVICES shards the blocks contiguously over one engine per listed
    device (a device may be listed more than once, so the sharded path also runs on a one-GPU box); every shard is
    driven by its own host thread.  Same objective, same master columns and the same relaxed_solution as one engine."""
    L = lib()
    prob = P.gen_mcf(40, 16, 40, 8, seed=3)
    paths = write_example(prob, tmp_path)
    subs = (C.c_char_p * prob.K)(*[s.encode() for s in paths["subs"]])
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        def run():
            o = DwOptions()
            L.dw_options_init(C.byref(o))
            o.verbosity = -1
            res = DwResult()
            assert L.dw_solve(paths["master"].encode(), subs, prob.K, C.byref(o), C.byref(res)) == 0
            out = (res.objective_value, res.num_vars, sorted(open("relaxed_solution").read().splitlines()))
            L.dw_result_free(C.byref(res))
            os.remove("relaxed_solution")
            return out
        monkeypatch.delenv("DWSOLVER_DEVICES", raising=False)
        one = run()
        monkeypatch.setenv("DWSOLVER_DEVICES", devices)
        many = run()
        assert abs(many[0] - one[0]) <= 1e-9 * (1 + abs(one[0]))
        assert many[1] == one[1] and many[2] == one[2]
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_dw_solve_sharded_example_with_fewer_blocks_than_devices(golden, tmp_path, monkeypatch):
    """more listed devices than blocks: the list is cut to K; the reference's book_dantzig objective is reached"""
    monkeypatch.setenv("DWSOLVER_DEVICES", "0,0,0,0,0,0,0,0")
    prob = load_example("book_dantzig")
    from dwsolver_b200.solver import dw_solve
    paths = write_example(prob, tmp_path)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        r = dw_solve(paths["master"], paths["subs"], verbosity=-1, shift=float(prob.shift or 0.0))
    finally:
        os.chdir(cwd)
    want = golden["objective"]["book_dantzig"]
    assert r["status"] == 0 and abs(r["objective"] - want) <= 1e-7 * (1 + abs(want))


def _gub_case(rng, L, K):
    """a random DW-shaped master: (cols, cost, rhs, lam) — L linking equalities, K convexity rows, 3-6 columns per
    block, one +1 slack and one costly -1 column per linking row"""
    cols, cost, x0 = [], [], []
    for k in range(K):
        nk = int(rng.integers(3, 7))
        wts = rng.dirichlet(np.ones(nk))
        for t in range(nk):
            a = np.where(rng.random(L) < min(0.5, 6.0 / L), rng.integers(-3, 4, L), 0).astype(float)
            cols.append((list(np.nonzero(a)[0]) + [L + k], list(a[np.nonzero(a)[0]]) + [1.0]))
            cost.append(float(rng.integers(-5, 6)))
            x0.append(wts[t])
    lam = len(cols)
    act = np.zeros(L)
    for j, (ix, vv) in enumerate(cols):
        for i, v in zip(ix, vv):
            if i < L:
                act[i] += v * x0[j]
    rhs = np.concatenate([act + rng.integers(0, 3, L), np.ones(K)])
    for i in range(L):
        cols.append(([i], [1.0])); cost.append(0.0)
        cols.append(([i], [-1.0])); cost.append(7.0)
    return cols, np.array(cost), rhs, lam


@pytest.mark.parametrize("block", [0, 8])
def test_gub_master_thread_team_is_deterministic(block, monkeypatch):
    """(block = 8: the same with blocked updates of the working inverse, DWSOLVER_MASTER_BLOCK — pending rank-1 pairs
    folded into FTRAN, duals and pivot rows, applied in one pass every 8 column replacements.)
    L = 144 linking rows: large enough for the master's thread team (host/thread_team.h).  One, three and eight
    members must walk the SAME simplex path — equal iteration counts and bit-equal primal and dual vectors — and
    reach the HiGHS optimum."""
    L_ = lib()
    rng = np.random.default_rng(4711)
    L, K = 144, 40
    cols, cost, rhs, lam = _gub_case(rng, L, K)
    n, m = len(cols), L + K
    A = sp.lil_matrix((m, n))
    for j, (ix, vv) in enumerate(cols):
        for i, v in zip(ix, vv):
            A[i, j] = v
    A = A.tocsc()
    ref = linprog(cost, A_eq=A, b_eq=rhs, bounds=[(0, None)] * n, method="highs")
    assert ref.status == 0
    colptr = np.zeros(n + 1, np.int32)
    idx, val = [], []
    for j, (ix, vv) in enumerate(cols):
        idx += [int(i) for i in ix]; val += [float(v) for v in vv]
        colptr[j + 1] = len(idx)
    idx = np.array(idx, np.int32); val = np.array(val)
    seen = []
    monkeypatch.setenv("DWSOLVER_MASTER_BLOCK", str(block))
    for threads in (1, 3, 8):
        monkeypatch.setenv("DWSOLVER_MASTER_THREADS", str(threads))
        x, y = np.zeros(n), np.zeros(m)
        obj, its = C.c_double(0), C.c_longlong(0)
        rc = hooks().dwhost_gub_solve(L, K, n, colptr.ctypes.data_as(c_ip), idx.ctypes.data_as(c_ip), val.ctypes.data_as(c_dp),
                                 cost.ctypes.data_as(c_dp), rhs.ctypes.data_as(c_dp), 0, x.ctypes.data_as(c_dp),
                                 y.ctypes.data_as(c_dp), C.byref(obj), C.byref(its))
        assert rc == 0
        assert abs(obj.value - ref.fun) <= 1e-8 * (1 + abs(ref.fun)), (threads, obj.value, ref.fun)
        assert np.abs(A @ x - rhs).max() <= 1e-8 and x.min() >= -1e-9
        seen.append((its.value, x.copy(), y.copy()))
    assert seen[0][0] > L                                   # a real solve, not a crash basis that happened to be optimal
    for other in seen[1:]:
        assert other[0] == seen[0][0] and np.array_equal(other[1], seen[0][1]) and np.array_equal(other[2], seen[0][2])


@pytest.mark.parametrize("seed", range(8))
def test_gub_master_with_blocked_updates_matches_highs(seed, monkeypatch):
    monkeypatch.setenv("DWSOLVER_MASTER_BLOCK", str(2 + 3 * (seed % 4)))       # k = 2, 5, 8, 11
    test_gub_master_matches_highs(10 + 6 * seed)                               # L = 14 ... 56


@pytest.mark.parametrize("seed", range(8))
def test_gub_master_matches_highs(seed):
    """random DW-shaped masters: L linking equalities with slack/surplus columns, K convexity rows, several columns
    per block; solved cold and with a warm second stage (columns added to an optimal basis)"""
    L_ = lib()
    rng = np.random.default_rng(100 + seed)
    L, K = 4 + seed, 3 + 2 * seed
    cols, cost = [], []
    x0 = []
    for k in range(K):                                  # 3-6 columns per block; the first carries weight in x0
        nk = int(rng.integers(3, 7))
        wts = rng.dirichlet(np.ones(nk))
        for t in range(nk):
            a = np.where(rng.random(L) < 0.5, rng.integers(-3, 4, L), 0).astype(float)
            cols.append((list(np.nonzero(a)[0]) + [L + k], list(a[np.nonzero(a)[0]]) + [1.0]))
            cost.append(float(rng.integers(-5, 6)))
            x0.append(wts[t])
    lam = len(cols)
    A_link = np.zeros((L, lam))
    for j, (ix, vv) in enumerate(cols):
        for i, v in zip(ix, vv):
            if i < L:
                A_link[i, j] = v
    act = A_link @ np.array(x0)
    rhs = np.concatenate([act + rng.integers(0, 3, L), np.ones(K)])      # slack needed on some rows
    for i in range(L):                                   # one +1 slack and one -1 artificial-like column per row
        cols.append(([i], [1.0])); cost.append(0.0)
        cols.append(([i], [-1.0])); cost.append(7.0)
    n = len(cols)
    m = L + K
    A = sp.lil_matrix((m, n))
    for j, (ix, vv) in enumerate(cols):
        for i, v in zip(ix, vv):
            A[i, j] = v
    A = A.tocsc()
    cost = np.array(cost)
    ref = linprog(cost, A_eq=A, b_eq=rhs, bounds=[(0, None)] * n, method="highs")
    assert ref.status == 0
    colptr = np.zeros(n + 1, np.int32)
    idx, val = [], []
    for j, (ix, vv) in enumerate(cols):
        idx += [int(i) for i in ix]; val += [float(v) for v in vv]
        colptr[j + 1] = len(idx)
    idx = np.array(idx, np.int32); val = np.array(val)
    for n_first in (0, lam // 2 + 2 * L):                # (second variant: not all blocks have a column at first -> skipped)
        if n_first:
            have = {ix[-1] - L for (ix, vv) in cols[:n_first] if ix[-1] >= L}
            if len(have) < K:
                continue
        x, y = np.zeros(n), np.zeros(m)
        obj, its = C.c_double(0), C.c_longlong(0)
        rc = hooks().dwhost_gub_solve(L, K, n, colptr.ctypes.data_as(c_ip), idx.ctypes.data_as(c_ip), val.ctypes.data_as(c_dp),
                                 cost.ctypes.data_as(c_dp), rhs.ctypes.data_as(c_dp), n_first, x.ctypes.data_as(c_dp),
                                 y.ctypes.data_as(c_dp), C.byref(obj), C.byref(its))
        assert rc == 0
        assert abs(obj.value - ref.fun) <= 1e-8 * (1 + abs(ref.fun)), (obj.value, ref.fun)
        assert np.abs(A @ x - rhs).max() <= 1e-8 and x.min() >= -1e-9
        d = cost - A.T @ y
        assert d.min() >= -1e-7                            # dual feasibility of (pi, mu)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["book_dantzig", "four_sea", "web_trick", "book_lasdon"])
def test_dw_solve_blocks_matches_goldens(name, golden, tmp_path, monkeypatch):
    """the in-memory entry point (no LP text) reaches the same optimum as the reference's test-suite pins"""
    from dwsolver_b200.solver import dw_solve_blocks
    monkeypatch.chdir(tmp_path)
    prob = load_example(name)
    out = dw_solve_blocks(prob, verbosity=-1)
    want = golden["objective"][name]
    assert out["status"] == 0 and abs(out["objective"] - want) <= 1e-7 * (1 + abs(want))
    assert out["stats"]["dw_iterations"] > 0 and os.path.exists(tmp_path / "relaxed_solution")


@pytest.mark.gpu
def test_dw_solve_blocks_config_a_both_masters(monkeypatch, tmp_path):
    """config A (1024 blocks) through dw_solve_blocks with the GUB master and with the dense-inverse master"""
    from dwsolver_b200.solver import dw_solve_blocks
    monkeypatch.chdir(tmp_path)
    prob = P.make_config("A")
    z = np.load(os.path.join(ROOT, "tests", "golden", "traj_A.npz"))
    want = float(z["objective"])
    a = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    assert a["status"] == 0 and abs(a["objective"] - want) <= 1e-7 * (1 + abs(want))
    monkeypatch.setenv("DWSOLVER_MASTER", "dense")
    b = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
    assert b["status"] == 0 and abs(b["objective"] - want) <= 1e-7 * (1 + abs(want))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mcf", "examples", "infeasible"])
def test_dw_solve_with_the_device_master(case, monkeypatch, tmp_path, golden):
    """DWSOLVER_MASTER=ipm: the restricted master solved by the device-resident interior-point method
    (csrc/master_ipm.cu behind host/ipm_master.cc) through the whole drop-in loop — Phase I with artificials, Phase II,
    reconstruction from the (non-vertex) master solution.  Same optimum as the simplex master to 1e-7, and the
    reconstructed x written to relaxed_solution is feasible for every block and every linking row and attains it."""
    from dwsolver_b200.solver import dw_solve_blocks
    monkeypatch.chdir(tmp_path)
    if case == "infeasible":
        # x1 >= 1, x2 >= 1, x1 + x2 <= 1: the device master cannot converge on an infeasible Phase II master; the simplex
        # master takes over and reports it
        monkeypatch.setenv("DWSOLVER_MASTER", "ipm")
        inf = P.BlockAngularLP(K=2, L=1, blocks=[_tiny_block(1.0, 5.0, 1.0), _tiny_block(1.0, 5.0, 1.0)],
                               link_lb=np.array([-np.inf]), link_ub=np.array([1.0]))
        assert dw_solve_blocks(inf, verbosity=-1, print_relaxed_sol=0)["status"] == 3
        return
    probs = [P.gen_mcf(96, 64, 128, 32, 1024001)] if case == "mcf" else [load_example(n) for n in ("book_dantzig", "four_sea", "web_trick", "book_lasdon")]
    for prob in probs:
        monkeypatch.setenv("DWSOLVER_MASTER", "gub")
        ref = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
        monkeypatch.setenv("DWSOLVER_MASTER", "ipm")
        out = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=1)
        assert ref["status"] == 0 and out["status"] == 0
        assert abs(out["objective"] - ref["objective"]) <= 1e-7 * (1 + abs(ref["objective"])), (out["objective"], ref["objective"])
        # the reconstruction: relaxed_solution names the variables x_<k>_<j>
        xs = [np.zeros(b.n) for b in prob.blocks]
        for line in open(tmp_path / "relaxed_solution"):
            nm, v = line.split()
            _, k, j = nm.split("_")
            xs[int(k)][int(j)] = float(v)
        link = np.zeros(prob.L)
        total = float(getattr(prob, "shift", 0.0) or 0.0)
        for b, x in zip(prob.blocks, xs):
            act = sp.csr_matrix((b.a_val, b.a_col, b.a_rowptr), shape=(b.m, b.n)) @ x
            ub = P.effective_col_ub(b)
            tol = 2e-5          # relaxed_solution prints %f (six decimals)
            assert (act >= b.row_lb - tol * (1 + np.abs(act))).all() and (act <= b.row_ub + tol * (1 + np.abs(act))).all()
            assert (x >= b.col_lb - tol).all() and (x <= ub + tol).all()
            np.add.at(link, b.d_row, x[b.d_col] * b.d_val)
            total += float(b.c_master @ x)
        assert (link >= prob.link_lb - 1e-3 * (1 + np.abs(link))).all() and (link <= prob.link_ub + 1e-3 * (1 + np.abs(link))).all()
        assert abs(total - out["objective"]) <= 1e-4 * (1 + abs(out["objective"]))


def _tiny_block(lb, ub, c, d_rows=(0,), d_vals=(1.0,)):
    z = np.zeros(0)
    return P.Block(m=1, n=1, a_rowptr=np.array([0, 1], np.int32), a_col=np.array([0], np.int32), a_val=np.array([1.0]),
                   row_lb=np.array([-np.inf]), row_ub=np.array([np.inf]), col_lb=np.array([lb]), col_ub=np.array([ub]),
                   c_file=np.array([c]), c_master=np.array([c]), d_row=np.array(d_rows, np.int32),
                   d_col=np.zeros(len(d_rows), np.int32), d_val=np.array(d_vals, float))


@pytest.mark.gpu
def test_dw_solve_blocks_status_codes(monkeypatch, tmp_path):
    """infeasible linking row -> DW_STATUS_ERR_INFEASIBLE (3); a block unbounded below -> DW_STATUS_ERR_UNBOUNDED (6)
    (src/dw_solver.h:88-97; the reference prints and carries on, this build returns the code)"""
    from dwsolver_b200.solver import dw_solve_blocks
    monkeypatch.chdir(tmp_path)
    # x1 >= 1, x2 >= 1, x1 + x2 <= 1
    inf = P.BlockAngularLP(K=2, L=1, blocks=[_tiny_block(1.0, 5.0, 1.0), _tiny_block(1.0, 5.0, 1.0)],
                           link_lb=np.array([-np.inf]), link_ub=np.array([1.0]))
    out = dw_solve_blocks(inf, verbosity=-1, print_relaxed_sol=0)
    assert out["status"] == 3
    # feasible twin: x1 + x2 <= 3 -> optimum 2 at (1, 1)
    ok = P.BlockAngularLP(K=2, L=1, blocks=[_tiny_block(1.0, 5.0, 1.0), _tiny_block(1.0, 5.0, 1.0)],
                          link_lb=np.array([-np.inf]), link_ub=np.array([3.0]))
    out = dw_solve_blocks(ok, verbosity=-1, print_relaxed_sol=0)
    assert out["status"] == 0 and abs(out["objective"] - 2.0) <= 1e-9
    # min x with x free below: unbounded subproblem
    unb = P.BlockAngularLP(K=1, L=1, blocks=[_tiny_block(-np.inf, 5.0, 1.0)], link_lb=np.array([-np.inf]),
                           link_ub=np.array([3.0]))
    out = dw_solve_blocks(unb, verbosity=-1, print_relaxed_sol=0)
    assert out["status"] == 6


@pytest.mark.gpu
@pytest.mark.parametrize("master", ["auto", "ipm"])
def test_random_instances_match_highs(master, monkeypatch, tmp_path):
    """twelve random block-angular instances (network and dense blocks, all three small-block kernel paths) through
    dw_solve_blocks against the HiGHS optimum of the monolithic LP (profiles/stress_dw.py runs more of them) — with the
    master the library picks by size (the host simplex at these sizes) and with the device interior-point master forced"""
    from dwsolver_b200.solver import dw_solve_blocks
    from highs_util import solve_monolithic
    monkeypatch.chdir(tmp_path)
    if master == "ipm":
        monkeypatch.setenv("DWSOLVER_MASTER", "ipm")
    rng = np.random.default_rng(77)
    for t in range(12):
        if t % 3 == 0:
            nodes = int(rng.choice([16, 32, 64]))
            prob = P.gen_mcf(int(rng.integers(8, 120)), nodes, nodes * 2 + int(rng.integers(0, 20)), int(rng.integers(2, nodes // 2)),
                             seed=int(rng.integers(1, 10**6)))
        elif t % 3 == 1:
            prob = P.gen_mcf(int(rng.integers(8, 40)), 256, 512, int(rng.integers(8, 64)), seed=int(rng.integers(1, 10**6)))
        else:
            m = int(rng.integers(20, 100))
            prob = P.gen_dense(int(rng.integers(2, 8)), m, 2 * m + int(rng.integers(0, 40)), int(rng.integers(2, 12)),
                               seed=int(rng.integers(1, 10**6)), dens_a=0.08)
        out = dw_solve_blocks(prob, verbosity=-1, print_relaxed_sol=0)
        ref = solve_monolithic(prob)["objective"]
        assert out["status"] == 0 and abs(out["objective"] - ref) <= 1e-7 * (1 + abs(ref)), (t, out["status"], out["objective"], ref)
