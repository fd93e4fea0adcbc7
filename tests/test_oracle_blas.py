"""Pin the oracle's dw_blas restatement (oracle/orc_blas.c) against
 (1) the known-answer vectors of the reference's unit test (tests/test_blas.c:47-184,
     tolerance 1e-12 at :42) committed in tests/golden/golden.json, and
 (2) the reference's own dw_blas.c compiled into oracle/_ref/libdwblas_ref.so, bit for bit,
     on seeded random inputs (skipped when the prebuilt reference object is absent)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

TOL = 1e-12


def d(a):
    return a.ctypes.data_as(O.c_dp)


def i(a):
    return a.ctypes.data_as(O.c_ip)


def f64(v):
    return np.asarray(v, dtype=np.float64).copy()


def i32(v):
    return np.asarray(v, dtype=np.int32).copy()


def test_known_answers(golden):
    L = O.lib()
    g = golden["blas"]
    for case in g["daxpy"]:
        x, y = f64(case["x"] or [99.0]), f64(case["y"] or [7.0])
        keep = y.copy()
        L.orc_daxpy(len(case["x"]), case["alpha"], d(x), d(y))
        exp = f64(case["out"]) if case["out"] else keep
        assert np.allclose(y[:len(exp)], exp, atol=TOL, rtol=0)
    for case in g["ddot"]:
        x, y = f64(case["x"] or [1.0]), f64(case["y"] or [1.0])
        assert abs(L.orc_ddot(len(case["x"]), d(x), d(y)) - case["out"]) <= TOL
    for case in g["dcopy"]:
        x = f64(case["x"] or [42.0])
        y = np.full(max(len(case["x"]), 1), 99.0)
        L.orc_dcopy(len(case["x"]), d(x), d(y))
        if case["x"]:
            assert np.array_equal(y, x)
        else:
            assert y[0] == 99.0
    for case in g["ddoti"]:
        x, ind, y = f64(case["x"] or [2.0]), i32(case["ind"] or [0]), f64(case["y"])
        assert abs(L.orc_ddoti(len(case["x"]), d(x), i(ind), d(y)) - case["out"]) <= TOL
    for case in g["dcoogemv"]:
        val, ix, jx = f64(case["val"] or [1.0]), i32(case["indx"] or [0]), i32(case["jndx"] or [0])
        b = f64(case["b"])
        c = np.full(case["m"], 99.0)
        L.orc_dcoogemv(case["m"], case["k"], d(val), i(ix), i(jx), len(case["val"]), d(b), d(c))
        assert np.allclose(c, f64(case["out"]), atol=TOL, rtol=0)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bitwise_vs_reference_object(seed):
    R = O.ref_blas()
    if R is None:
        pytest.skip("oracle/_ref/libdwblas_ref.so not built (reference tree absent)")
    L = O.lib()
    rng = np.random.default_rng(seed)
    n = 257
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    y1, y2 = y.copy(), y.copy()
    L.orc_daxpy(n, -1.0, d(x), d(y1)); R.dw_daxpy(n, -1.0, d(x), d(y2))
    assert np.array_equal(y1, y2)
    L.orc_daxpy(n, 0.3, d(x), d(y1)); R.dw_daxpy(n, 0.3, d(x), d(y2))
    assert np.array_equal(y1, y2)
    assert L.orc_ddot(n, d(x), d(y)) == R.dw_ddot(n, d(x), d(y))
    ind = rng.integers(0, n, 64).astype(np.int32)
    assert L.orc_ddoti(64, d(x), i(ind), d(y)) == R.dw_ddoti(64, d(x), i(ind), d(y))
    L.orc_dcopy(n, d(x), d(y1)); R.dw_dcopy(n, d(x), d(y2))
    assert np.array_equal(y1, y2)
    m, k, nnz = 40, 97, 900
    val = rng.standard_normal(nnz)
    ix = rng.integers(0, m, nnz).astype(np.int32)
    jx = rng.integers(0, k, nnz).astype(np.int32)
    b = rng.standard_normal(k)
    c1, c2 = np.full(m, 3.0), np.full(m, -3.0)
    L.orc_dcoogemv(m, k, d(val), i(ix), i(jx), nnz, d(b), d(c1))
    R.dw_dcoogemv(m, k, d(val), i(ix), i(jx), nnz, d(b), d(c2))
    assert np.array_equal(c1, c2)
