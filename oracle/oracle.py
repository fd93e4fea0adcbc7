"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (see oracle/oracle.h).  The product path (dwsolver_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_blas.c", "orc_simplex.c", "orc_worker.c", "oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/dw_blas.c") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libdwblas_ref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


class Proposal(C.Structure):
    _fields_ = [("status", C.c_int), ("obj", C.c_double), ("new_obj_coeff", C.c_double),
                ("len", C.c_int), ("ind", c_ip), ("val", c_dp), ("x", c_dp)]


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_daxpy.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.orc_ddot.argtypes = [C.c_int, c_dp, c_dp]; L.orc_ddot.restype = C.c_double
        L.orc_ddoti.argtypes = [C.c_int, c_dp, c_ip, c_dp]; L.orc_ddoti.restype = C.c_double
        L.orc_dcopy.argtypes = [C.c_int, c_dp, c_dp]
        L.orc_dcoogemv.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, C.c_int, c_dp, c_dp]
        L.orc_lp_new.restype = C.c_void_p
        L.orc_lp_new.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.orc_lp_free.argtypes = [C.c_void_p]
        L.orc_lp_solve.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]; L.orc_lp_solve.restype = C.c_int
        L.orc_lp_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_lp_residual.argtypes = [C.c_void_p, c_dp]; L.orc_lp_residual.restype = C.c_double
        L.orc_block_new.restype = C.c_void_p
        L.orc_block_new.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                    C.c_int, C.c_int, c_ip, c_ip, c_dp]
        L.orc_block_free.argtypes = [C.c_void_p]
        L.orc_block_initial.argtypes = [C.c_void_p, C.POINTER(Proposal)]
        L.orc_block_iterate.argtypes = [C.c_void_p, c_dp, C.c_int, C.POINTER(Proposal)]
        L.orc_block_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_batch.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Proposal), C.c_int, c_dp, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def ref_blas():
    """The reference's own dw_blas.c compiled into oracle/_ref (None when not built)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libdwblas_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.dw_daxpy.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        R.dw_ddot.argtypes = [C.c_int, c_dp, c_dp]; R.dw_ddot.restype = C.c_double
        R.dw_ddoti.argtypes = [C.c_int, c_dp, c_ip, c_dp]; R.dw_ddoti.restype = C.c_double
        R.dw_dcopy.argtypes = [C.c_int, c_dp, c_dp]
        R.dw_dcoogemv.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, C.c_int, c_dp, c_dp]
        _REF = R
    return _REF


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _bounds(a):
    a = _f64(a).copy()
    a[np.isposinf(a)] = 1e30
    a[np.isneginf(a)] = -1e30
    return a


class OracleLP:
    """One LP with a persistent (warm-startable) basis."""

    def __init__(self, m, n, rowptr, colidx, vals, row_lb, row_ub, col_lb, col_ub):
        self.m, self.n = int(m), int(n)
        self._keep = [_i32(rowptr), _i32(colidx), _f64(vals), _bounds(row_lb), _bounds(row_ub),
                      _bounds(col_lb), _bounds(col_ub)]
        k = self._keep
        self._h = lib().orc_lp_new(self.m, self.n, _i(k[0]), _i(k[1]), _d(k[2]), _d(k[3]), _d(k[4]), _d(k[5]), _d(k[6]))

    def solve(self, cost):
        cost = _f64(cost)
        x = np.zeros(max(self.n, 1))
        z = C.c_double(0.0)
        st = lib().orc_lp_solve(self._h, _d(cost), _d(x), C.byref(z))
        return st, z.value, x[:self.n]

    def residual(self, x):
        x = _f64(x)
        return lib().orc_lp_residual(self._h, _d(x))

    def counters(self):
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        lib().orc_lp_counters(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_lp_free(self._h)
            self._h = None


class OracleBlocks:
    """All K block workers of a BlockAngularLP (dwsolver_b200.problem.BlockAngularLP)."""

    def __init__(self, prob):
        self.prob = prob
        self.K, self.L = prob.K, prob.L
        self._keep = []
        self._h = (C.c_void_p * self.K)()
        self._out = (Proposal * self.K)()
        self.x = []
        self.ind = []
        self.val = []
        for k, b in enumerate(prob.blocks):
            arrs = [_i32(b.a_rowptr), _i32(b.a_col), _f64(b.a_val), _bounds(b.row_lb), _bounds(b.row_ub),
                    _bounds(b.col_lb), _bounds(b.col_ub), _f64(b.c_file), _f64(b.c_master),
                    _i32(b.d_row), _i32(b.d_col), _f64(b.d_val)]
            self._keep.append(arrs)
            self._h[k] = lib().orc_block_new(b.m, b.n, _i(arrs[0]), _i(arrs[1]), _d(arrs[2]), _d(arrs[3]), _d(arrs[4]),
                                             _d(arrs[5]), _d(arrs[6]), _d(arrs[7]), _d(arrs[8]),
                                             self.L, len(arrs[11]), _i(arrs[9]), _i(arrs[10]), _d(arrs[11]))
            x = np.zeros(max(b.n, 1)); ind = np.zeros(max(self.L, 1), np.int32); val = np.zeros(max(self.L, 1))
            self.x.append(x); self.ind.append(ind); self.val.append(val)
            self._out[k].x = _d(x); self._out[k].ind = _i(ind); self._out[k].val = _d(val)

    def _collect(self):
        out = []
        for k in range(self.K):
            o = self._out[k]
            n = self.prob.blocks[k].n
            out.append(dict(status=o.status, obj=o.obj, cost=o.new_obj_coeff, len=o.len,
                            ind=self.ind[k][:o.len].copy(), val=self.val[k][:o.len].copy(),
                            x=self.x[k][:n].copy()))
        return out

    def initial(self, nthreads: int = 1):
        lib().orc_batch(self._h, self._out, self.K, None, 0, 1, nthreads)
        return self._collect()

    def iterate(self, pi, phase_one: bool, nthreads: int = 1, collect: bool = True):
        pi = _f64(pi)
        lib().orc_batch(self._h, self._out, self.K, _d(pi), int(bool(phase_one)), 0, nthreads)
        return self._collect() if collect else None

    def counters(self):
        tot = np.zeros(3, np.int64)
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        for k in range(self.K):
            lib().orc_block_counters(self._h[k], C.byref(a), C.byref(b), C.byref(c))
            tot += (a.value, b.value, c.value)
        return tuple(int(v) for v in tot)

    def __del__(self):
        for k in range(getattr(self, "K", 0)):
            if self._h[k]:
                lib().orc_block_free(self._h[k])
                self._h[k] = None
