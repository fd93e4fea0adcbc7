"""ctypes mirror of the drop-in host library (include/dwsolver.h -> dwsolver_b200/lib/libdwsolver_b200.so):
the reference's public API dw_solve (LP files), this build's dw_solve_blocks (in-memory instance) and the binary
container pair dw_write_container / dw_solve_container."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import engine as E

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdwsolver_b200.so")
c_dp = C.POINTER(C.c_double)


class DwOptions(C.Structure):          # include/dwsolver.h, same order as the reference's src/dw_solver.h:105-118
    _fields_ = [("verbosity", C.c_int), ("mip_gap", C.c_double), ("max_phase1_iterations", C.c_int),
                ("max_phase2_iterations", C.c_int), ("rounding_flag", C.c_int), ("integerize_flag", C.c_int),
                ("enforce_sub_integrality", C.c_int), ("print_timing_data", C.c_int), ("print_final_master", C.c_int),
                ("print_relaxed_sol", C.c_int), ("perturb", C.c_int), ("shift", C.c_double)]


class DwResult(C.Structure):
    _fields_ = [("status", C.c_int), ("objective_value", C.c_double), ("num_vars", C.c_int), ("x", c_dp)]


class DwStats(C.Structure):
    _fields_ = [("dw_iterations", C.c_int), ("phase1_iterations", C.c_int), ("master_columns", C.c_int),
                ("sub_pivots", C.c_longlong), ("seconds_total", C.c_double), ("seconds_read", C.c_double),
                ("seconds_subproblems", C.c_double), ("seconds_master", C.c_double)]


_LIB = None


def load_library():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — build it with `make -C dwsolver_b200/host`")
        L = C.CDLL(LIB_PATH)
        L.dw_version.restype = C.c_char_p
        L.dw_solve.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(DwOptions), C.POINTER(DwResult)]
        L.dw_solve_blocks.argtypes = [C.c_int, C.c_int, C.POINTER(E.BlockDesc), c_dp, c_dp, C.POINTER(DwOptions),
                                      C.POINTER(DwResult)]
        L.dw_write_container.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_double, C.c_char_p]
        L.dw_solve_container.argtypes = [C.c_char_p, C.POINTER(DwOptions), C.POINTER(DwResult)]
        _LIB = L
    return _LIB


def _options(**kw) -> DwOptions:
    o = DwOptions()
    load_library().dw_options_init(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown dw_options_t field {k!r}")
        setattr(o, k, v)
    return o


def _result(lib, rc: int, res: DwResult) -> dict:
    st = DwStats()
    lib.dw_last_stats(C.byref(st))
    out = dict(status=rc, objective=res.objective_value,
               x=np.ctypeslib.as_array(res.x, (res.num_vars,)).copy() if res.x and res.num_vars else np.zeros(0),
               stats={f: getattr(st, f) for f, _ in DwStats._fields_})
    lib.dw_result_free(C.byref(res))
    return out


def dw_solve(master_file: str, subproblem_files: Sequence[str], **opts) -> dict:
    lib = load_library()
    subs = (C.c_char_p * len(subproblem_files))(*[s.encode() for s in subproblem_files])
    res = DwResult()
    rc = lib.dw_solve(master_file.encode(), subs, len(subproblem_files), C.byref(_options(**opts)), C.byref(res))
    return _result(lib, rc, res)


def dw_write_container(master_file: str, subproblem_files: Sequence[str], container_path: str, shift: float = 0.0) -> int:
    """Parse the LP files once and store the assembled instance (include/dwsolver.h); returns the status code."""
    lib = load_library()
    subs = (C.c_char_p * len(subproblem_files))(*[s.encode() for s in subproblem_files])
    return lib.dw_write_container(master_file.encode(), subs, len(subproblem_files), float(shift), container_path.encode())


def dw_solve_container(container_path: str, **opts) -> dict:
    lib = load_library()
    res = DwResult()
    rc = lib.dw_solve_container(container_path.encode(), C.byref(_options(**opts)), C.byref(res))
    return _result(lib, rc, res)


def dw_solve_blocks(prob, **opts) -> dict:
    """dw_solve on a BlockAngularLP held in memory (no LP text)."""
    lib = load_library()
    keep = []
    descs = (E.BlockDesc * prob.K)()
    for k, b in enumerate(prob.blocks):
        arrs = dict(a_rowptr=E._i32(b.a_rowptr), a_col=E._i32(b.a_col), a_val=E._f64(b.a_val),
                    row_lb=E._bnd(b.row_lb), row_ub=E._bnd(b.row_ub), col_lb=E._bnd(b.col_lb), col_ub=E._bnd(b.col_ub),
                    c_file=E._f64(b.c_file), c_master=E._f64(b.c_master),
                    d_row=E._i32(b.d_row), d_col=E._i32(b.d_col), d_val=E._f64(b.d_val))
        keep.append(arrs)
        d = descs[k]
        d.m, d.n, d.d_nnz = int(b.m), int(b.n), int(len(arrs["d_val"]))
        for name, arr in arrs.items():
            setattr(d, name, arr.ctypes.data_as(E.c_ip if arr.dtype == np.int32 else E.c_dp))
    llb, lub = E._bnd(prob.link_lb), E._bnd(prob.link_ub)
    o = _options(**opts)
    if "shift" not in opts:
        o.shift = float(getattr(prob, "shift", 0.0) or 0.0)
    res = DwResult()
    rc = lib.dw_solve_blocks(prob.K, prob.L, descs, llb.ctypes.data_as(c_dp), lub.ctypes.data_as(c_dp), C.byref(o),
                             C.byref(res))
    return _result(lib, rc, res)
