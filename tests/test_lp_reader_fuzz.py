"""Differential test of the host library's CPLEX-LP reader (dwsolver_b200/host/lp_reader.cc, which replaces glp_read_lp:
/root/reference/src/dw_subprob.c:105-111, src/dw_solver.c:192-198) against the package's Python reader on randomly
FORMATTED LP text: every number style (integers, decimals, exponents, leading/trailing dots), repeated signs, line
breaks and comments inside expressions, all spellings of the section keywords and senses, every bound form (free,
ranges, one-sided, fixed, infinities, reversed), integer sections, names that begin like keywords or like exponents.
10 000 cases were run when the tokenizer was rewritten (no mismatch); the suite keeps 300."""
import ctypes as C
import os

import numpy as np

from conftest import ROOT
from dwsolver_b200 import problem as P

LIB = os.path.join(ROOT, "dwsolver_b200", "lib", "libdwsolver_b200.so")
HOOKS = os.path.join(ROOT, "dwsolver_b200", "lib", "libdwhost_testhooks.so")      # dwhost_* test entry points (tests/host_hooks)
c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def num(rng):
    k = rng.integers(0, 7)
    v = float(rng.integers(0, 2000)) / float(rng.choice([1, 1, 4, 8, 10, 1000]))
    if k == 0: return str(int(v))
    if k == 1: return repr(v)
    if k == 2: return "%.3e" % v
    if k == 3: return "%.6E" % v
    if k == 4: return ("%g" % v)
    if k == 5: return ("%.12f" % v).rstrip("0") + "0"
    return ".5" if rng.random() < 0.3 else "%d." % int(v)

def name(rng, i):
    return rng.choice(["x", "y_", "Var.", "s", "m", "b", "e", "st_", "inf_x", "free", "End_", "_z", "q!", "w#"]) + str(i)

def gen(rng):
    n = int(rng.integers(1, 9)); m = int(rng.integers(1, 6))
    names = []
    for i in range(n):
        nm = name(rng, i)
        if nm.lower() in ("free",): nm += "q"
        names.append(nm)
    ws = lambda: rng.choice([" ", "  ", "\n ", "\t", " \\ comment here\n "])
    def lin():
        parts = []
        for j in rng.permutation(n)[: int(rng.integers(1, n + 1))]:
            sign = rng.choice(["+", "-", "+ ", "- ", "+\n", " - -", "+ +"]) if parts or rng.random() < 0.5 else ""
            coef = "" if rng.random() < 0.3 else num(rng) + rng.choice([" ", ""," \n"])
            if coef and coef.strip()[-1] in "eE.": coef = coef.strip() + " "
            if coef and not coef.endswith((" ", "\n")) and names[j][0] in "eE": coef += " "
            parts.append(f"{sign}{ws()}{coef}{names[j]}")
        return ws().join(parts)
    txt = rng.choice(["Minimize", "minimize", "MIN", "Maximize", "maximum", "Minimum"]) + "\n"
    txt += rng.choice([" obj: ", " ", " cost:\n"]) + lin() + "\n"
    txt += rng.choice(["Subject To", "subject to", "st", "s.t.", "ST.", "such that"]) + "\n"
    for i in range(m):
        lab = rng.choice([f" r{i}: ", f" c.{i}:", " "])
        sense = rng.choice(["<=", ">=", "=", "=<", "=>", "<", ">"])
        rhs = rng.choice(["", "-", "+", "- "]) + num(rng)
        txt += lab + lin() + f" {sense} {rhs}\n"
    txt += rng.choice(["Bounds", "bounds", "BOUND"]) + "\n"
    for j in range(n):
        k = rng.integers(0, 8)
        a, b = num(rng), num(rng)
        if k == 0: txt += f" {names[j]} free\n"
        elif k == 1: txt += f" -{a} <= {names[j]} <= {b}\n"
        elif k == 2: txt += f" {names[j]} >= -{a}\n"
        elif k == 3: txt += f" {names[j]} <= {b}\n"
        elif k == 4: txt += f" {names[j]} = {a}\n"
        elif k == 5: txt += f" -inf <= {names[j]} <= +Infinity\n"
        elif k == 6: txt += f" {b} >= {names[j]}\n"
    if rng.random() < 0.4:
        txt += rng.choice(["General", "generals", "Integer", "Binary", "bin"]) + "\n " + " ".join(names[: int(rng.integers(1, n + 1))]) + "\n"
    txt += rng.choice(["End", "end", "END\n\n", ""]) + "\n"
    return txt


def read_with_host_library(L, path):
    dims = (C.c_int * 3)()
    if L.dwhost_lp_dims(path.encode(), dims):
        return None
    m, n, nnz = dims
    obj, lb, ub = np.zeros(n), np.zeros(n), np.zeros(n)
    rlb, rub = np.zeros(m), np.zeros(m)
    ar, ac, av = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
    names = C.create_string_buffer(64 * (n + 1))
    mx = C.c_int(0)
    d = lambda a: a.ctypes.data_as(c_dp)
    i = lambda a: a.ctypes.data_as(c_ip)
    assert L.dwhost_lp_read(path.encode(), d(obj), d(lb), d(ub), d(rlb), d(rub), i(ar), i(ac), d(av), names, len(names),
                            C.byref(mx)) == 0
    return dict(names=names.value.decode().split("\n")[:-1], obj=obj, lb=lb, ub=ub, rlb=rlb, rub=rub, ar=ar, ac=ac, av=av,
                maximize=bool(mx.value))


def test_lp_reader_agrees_with_the_python_reader_on_odd_formatting(tmp_path):
    L = C.CDLL(HOOKS)
    rng = np.random.default_rng(20261020)
    path = str(tmp_path / "f.lp")
    inf = lambda a: np.where(np.abs(a) >= 1e300, np.copysign(np.inf, a), a)
    for case in range(300):
        text = gen(rng)
        with open(path, "w") as fh:
            fh.write(text)
        ref = P.parse_lp_text(text)
        got = read_with_host_library(L, path)
        assert got is not None, text
        assert got["names"] == list(ref.col_names), text
        assert got["maximize"] == bool(ref.maximize), text
        for a, b in ((got["obj"], ref.obj), (got["lb"], ref.col_lb), (got["ub"], ref.col_ub), (got["rlb"], ref.row_lb),
                     (got["rub"], ref.row_ub), (got["av"], ref.a_vals)):
            assert np.array_equal(inf(np.asarray(a, float)), inf(np.asarray(b, float))), text
        assert np.array_equal(got["ar"], ref.a_rows) and np.array_equal(got["ac"], ref.a_cols), text
