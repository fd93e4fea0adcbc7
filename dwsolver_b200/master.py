"""ctypes mirror of the device-resident restricted master (include/dwgpu.h, dwgpu_master_*): a primal-dual
interior-point solve of  min c'x, A x = b, x >= 0  whose rows are L linking rows plus K convexity rows
(csrc/master_ipm.cu).  Replaces the reference's glp_simplex(master_lp) (src/dw_phases.c:219, 450) for large masters."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine as E

c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _lib():
    lib = E.load_library()
    if not getattr(lib, "_master_ready", False):
        lib.dwgpu_master_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.dwgpu_master_destroy.argtypes = [C.c_void_p]
        lib.dwgpu_master_destroy.restype = None
        lib.dwgpu_master_set_rhs.argtypes = [C.c_void_p, c_dp]
        lib.dwgpu_master_add_cols.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip, c_dp, c_dp]
        lib.dwgpu_master_set_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
        lib.dwgpu_master_set_coef.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        lib.dwgpu_master_del_cols.argtypes = [C.c_void_p, C.c_char_p]
        lib.dwgpu_master_cols.argtypes = [C.c_void_p]
        lib.dwgpu_master_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, c_ip]
        lib.dwgpu_master_get.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp]
        lib.dwgpu_master_quality.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
        lib.dwgpu_master_last_error.restype = C.c_char_p
        lib._master_ready = True
    return lib


class DeviceMaster:
    def __init__(self, L, K, device=0):
        self.lib = _lib()
        self.L, self.K = L, K
        self.h = C.c_void_p()
        rc = self.lib.dwgpu_master_create(device, L, K, C.byref(self.h))
        if rc:
            raise RuntimeError(f"dwgpu_master_create: {self.lib.dwgpu_master_last_error().decode()}")

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: {self.lib.dwgpu_master_last_error().decode()}")

    def set_rhs(self, b):
        b = np.ascontiguousarray(b, np.float64)
        assert len(b) == self.L + self.K
        self._check(self.lib.dwgpu_master_set_rhs(self.h, b.ctypes.data_as(c_dp)), "set_rhs")

    def add_cols(self, cols, cost):
        """cols: list of (row indices, values) over all L+K rows."""
        colptr = np.zeros(len(cols) + 1, np.int32)
        for j, (ix, _) in enumerate(cols):
            colptr[j + 1] = colptr[j] + len(ix)
        rows = np.ascontiguousarray(np.concatenate([np.asarray(ix, np.int32) for ix, _ in cols]) if cols else np.zeros(0), np.int32)
        vals = np.ascontiguousarray(np.concatenate([np.asarray(v, np.float64) for _, v in cols]) if cols else np.zeros(0), np.float64)
        cost = np.ascontiguousarray(cost, np.float64)
        self._check(self.lib.dwgpu_master_add_cols(self.h, len(cols), colptr.ctypes.data_as(c_ip), rows.ctypes.data_as(c_ip),
                                                    vals.ctypes.data_as(c_dp), cost.ctypes.data_as(c_dp)), "add_cols")

    def set_costs(self, first, cost):
        cost = np.ascontiguousarray(cost, np.float64)
        self._check(self.lib.dwgpu_master_set_costs(self.h, first, len(cost), cost.ctypes.data_as(c_dp)), "set_costs")

    def del_cols(self, flags):
        self._check(self.lib.dwgpu_master_del_cols(self.h, bytes(1 if f else 0 for f in flags)), "del_cols")

    def cols(self):
        return self.lib.dwgpu_master_cols(self.h)

    def solve(self, tol=1e-9, max_iter=100):
        st = C.c_int(0)
        self._check(self.lib.dwgpu_master_solve(self.h, tol, max_iter, C.byref(st)), "solve")
        n = self.cols()
        obj, dobj, its, secs = C.c_double(0), C.c_double(0), C.c_int(0), C.c_double(0)
        y, x = np.zeros(self.L + self.K), np.zeros(n)
        self._check(self.lib.dwgpu_master_get(self.h, C.byref(obj), C.byref(dobj), y.ctypes.data_as(c_dp), x.ctypes.data_as(c_dp),
                                              C.byref(its), C.byref(secs)), "get")
        q = [C.c_double(0) for _ in range(3)]
        self.lib.dwgpu_master_quality(self.h, *[C.byref(v) for v in q])
        return dict(status=st.value, pinf=q[0].value, dinf=q[1].value, gap=q[2].value, obj=obj.value, dual_obj=dobj.value, y=y, x=x, iterations=its.value, seconds=secs.value)

    def close(self):
        if self.h:
            self.lib.dwgpu_master_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
