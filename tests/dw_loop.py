"""Test-harness Dantzig-Wolfe loop (restates SURVEY.md Appendix A.2-A.6).

The restricted master is solved by scipy/HiGHS so that the loop is independent of the
product's own host master (dwsolver_b200/csrc/host).  The subproblem side is pluggable:
any object with ``initial()`` and ``iterate(pi, phase_one)`` returning per-block proposal
dicts (status, obj, cost, ind[1-based], val, x) — the oracle (oracle/oracle.py) or the
CUDA engine (dwsolver_b200/engine.py).

Reference: /root/reference/src/dw_solver.c:259-694 (master construction, phases),
src/dw_phases.c:54-304, 321-554 (accept test r_k - obj_k > 1e-6, column shape).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

TOLERANCE = 1e-6      # dw.h:71
DW_INFINITY = 1e6     # dw.h:72


class HighsMaster:
    """Restricted master: rows 0..L-1 linking (made FX), rows L..L+K-1 convexity (=1)."""

    def __init__(self, prob):
        self.L, self.K = prob.L, prob.K
        self.rhs = np.zeros(self.L + self.K)
        self.rhs[self.L:] = 1.0
        self.cols = []      # dict(name, cost, rows, vals, ub)
        lb, ub = prob.link_lb, prob.link_ub
        self.art = []
        for i in range(self.L):
            if lb[i] == ub[i]:                                  # FX -> artificial only (dw_solver.c:321-336)
                self.rhs[i] = ub[i]
                self._add(f"y_{i+1}", 1.0, [i], [-1.0 if ub[i] < 0 else 1.0], np.inf, art=True)
            elif np.isinf(lb[i]) and np.isfinite(ub[i]):        # UP -> surplus + artificial (338-375)
                self.rhs[i] = ub[i]
                self._add(f"su_{i+1}", 0.0, [i], [1.0], np.inf)
                self._add(f"y_{i+1}", 1.0, [i], [-1.0], np.inf, art=True)
            elif np.isinf(ub[i]) and np.isfinite(lb[i]):        # LO -> slack + artificial (381-419)
                self.rhs[i] = lb[i]
                self._add(f"sl_{i+1}", 0.0, [i], [-1.0], np.inf)
                self._add(f"y_{i+1}", 1.0, [i], [1.0], np.inf, art=True)
            else:
                raise ValueError("ranged/free linking rows are not supported by the reference (dw_solver.c:420-425)")

    def _add(self, name, cost, rows, vals, ub, art=False):
        self.cols.append(dict(name=name, cost=cost, rows=np.asarray(rows, np.int64),
                              vals=np.asarray(vals, np.float64), ub=ub, art=art, true_cost=cost))

    def add_lambda(self, k, it, cost, ind1, val, phase_one):
        rows = np.concatenate([np.asarray(ind1, np.int64) - 1, [self.L + k]])
        vals = np.concatenate([np.asarray(val, np.float64), [1.0]])
        self.cols.append(dict(name=f"lambda_{k}_{it}", cost=0.0 if phase_one else cost, rows=rows, vals=vals,
                              ub=1.0, art=False, true_cost=cost))

    def solve(self):
        n = len(self.cols)
        rows = np.concatenate([c["rows"] for c in self.cols])
        colsidx = np.concatenate([np.full(len(c["rows"]), j) for j, c in enumerate(self.cols)])
        vals = np.concatenate([c["vals"] for c in self.cols])
        A = sp.csr_matrix((vals, (rows, colsidx)), shape=(self.L + self.K, n))
        cost = np.array([c["cost"] for c in self.cols])
        bounds = [(0.0, None if np.isinf(c["ub"]) else c["ub"]) for c in self.cols]
        res = linprog(cost, A_eq=A, b_eq=self.rhs, bounds=bounds, method="highs-ds")
        if res.status != 0:
            return None
        duals = np.asarray(res.eqlin.marginals)
        return dict(obj=res.fun, x=res.x, pi=duals[:self.L].copy(), r=duals[self.L:].copy())


class ProductMaster(HighsMaster):
    """Same bookkeeping as HighsMaster, but the LP is solved by the PRODUCT's master solver (dwsolver_b200/host:
    GubMaster or the dense-inverse MasterLP) through its incremental test API, warm-started across rounds.
    Lets the whole DW loop run on the CPU (oracle subproblems) with the product's master."""

    def __init__(self, prob, gub=True):
        super().__init__(prob)
        import ctypes as C
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        self.C = C
        self.lib = C.CDLL(os.path.join(here, "..", "dwsolver_b200", "lib", "libdwhost_testhooks.so"))   # tests/host_hooks
        self.lib.dwhost_im_new.restype = C.c_void_p
        self.lib.dwhost_im_new.argtypes = [C.c_int, C.c_int, C.c_int]
        self.lib.dwhost_im_free.argtypes = [C.c_void_p]
        self.lib.dwhost_im_set_rhs.argtypes = [C.c_void_p, C.c_int, C.c_double]
        self.lib.dwhost_im_add_col.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        self.lib.dwhost_im_set_cost.argtypes = [C.c_void_p, C.c_int, C.c_double]
        self.lib.dwhost_im_del_cols.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        self.lib.dwhost_im_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        self.lib.dwhost_im_primal.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self.h = None
        self.gub = gub
        self.synced = []          # (dict object, cost sent)
        self.seconds = 0.0
        self.iterations = 0

    def _push(self, c):
        C = self.C
        idx = np.ascontiguousarray(c["rows"], np.int32)
        val = np.ascontiguousarray(c["vals"], np.float64)
        self.lib.dwhost_im_add_col(self.h, float(c["cost"]), 1e300 if np.isinf(c["ub"]) else float(c["ub"]), len(idx),
                                   idx.ctypes.data_as(C.POINTER(C.c_int)), val.ctypes.data_as(C.POINTER(C.c_double)))

    def solve(self):
        import time
        C = self.C
        if self.h is None:
            self.h = self.lib.dwhost_im_new(self.L, self.K, 1 if self.gub else 0)
            for i, b in enumerate(self.rhs):
                self.lib.dwhost_im_set_rhs(self.h, i, float(b))
        present = {id(c) for c in self.cols}
        if any(id(c) not in present for c, _ in self.synced):
            flags = bytes(0 if id(c) in present else 1 for c, _ in self.synced)
            self.lib.dwhost_im_del_cols(self.h, len(flags), flags)
            self.synced = [(c, cs) for c, cs in self.synced if id(c) in present]
        for j, (c, cs) in enumerate(self.synced):
            if c["cost"] != cs:
                self.lib.dwhost_im_set_cost(self.h, j, float(c["cost"]))
                self.synced[j] = (c, c["cost"])
        for c in self.cols[len(self.synced):]:
            self._push(c)
            self.synced.append((c, c["cost"]))
        obj, its = C.c_double(0), C.c_longlong(0)
        duals = np.zeros(self.L + self.K)
        t0 = time.perf_counter()
        rc = self.lib.dwhost_im_solve(self.h, C.byref(obj), duals.ctypes.data_as(C.POINTER(C.c_double)), C.byref(its))
        self.seconds += time.perf_counter() - t0
        self.iterations = its.value
        if rc != 0:
            return None
        x = np.zeros(len(self.cols))
        self.lib.dwhost_im_primal(self.h, x.ctypes.data_as(C.POINTER(C.c_double)))
        return dict(obj=obj.value, x=x, pi=duals[:self.L].copy(), r=duals[self.L:].copy())


def run_dw(prob, engine, max_p1=100, max_p2=3000, record=None, verbose=False, master=None):
    """Returns dict(objective, iterations, x (per-block reconstructed), trajectory)."""
    K, L = prob.K, prob.L
    master = master if master is not None else HighsMaster(prob)
    need_phase_one = any(c["art"] for c in master.cols)
    r = np.full(K, DW_INFINITY)
    X = {}                    # (k, iter) -> x
    it = 0
    traj = []

    def consume(props, phase_one):
        added = 0
        for k, p in enumerate(props):
            if r[k] - p["obj"] > TOLERANCE:
                master.add_lambda(k, it, p["cost"], p["ind"], p["val"], phase_one)
                X[(k, it)] = p["x"].copy()
                added += 1
        return added

    props = engine.initial()
    sol = None
    if need_phase_one:
        first = True
        for _ in range(max_p1):
            n_before = len(master.cols)
            added = consume(props, True)
            if first:                                           # dw_phases.c:165-171, 201-217
                acc = np.zeros(L)
                for c in master.cols[n_before:]:
                    rows, vals = c["rows"][:-1], c["vals"][:-1]
                    np.add.at(acc, rows, vals)
                for i in range(L):
                    if master.rhs[i] - acc[i] < 0.0:
                        for c in master.cols:
                            if c["art"] and c["name"] == f"y_{i+1}":
                                c["vals"] = np.array([-1.0])
                first = False
            sol = master.solve()
            if sol is None:
                raise RuntimeError("phase-1 master failed")
            r = sol["r"]
            it += 1
            if verbose:
                print(f"P1 it={it} added={added} obj={sol['obj']:.8e}")
            if abs(sol["obj"]) < TOLERANCE:
                break
            if added == 0:
                break
            if record is not None:
                record.append((sol["pi"].copy(), True))
            traj.append((sol["pi"].copy(), True))
            props = engine.iterate(sol["pi"], True)
        # leave phase one: drop artificials, restore costs (dw_solver.c:535-562)
        master.cols = [c for c in master.cols if not c["art"]]
        for c in master.cols:
            c["cost"] = c["true_cost"]
        sol = master.solve()
        if sol is None:
            return dict(objective=None, status="infeasible", iterations=it)
        r = sol["r"]
        it += 1
    else:
        added = consume(props, False)
        sol = master.solve()
        r = sol["r"]
        it += 1
    status = "phase2_limit"
    for _ in range(max_p2):
        traj.append((sol["pi"].copy(), False))
        if record is not None:
            record.append((sol["pi"].copy(), False))
        props = engine.iterate(sol["pi"], False)
        added = consume(props, False)
        if verbose:
            print(f"P2 it={it} added={added} obj={sol['obj'] + prob.shift:.10e}")
        if added == 0:
            status = "optimal"
            break
        sol = master.solve()
        r = sol["r"]
        it += 1
    # reconstruction (dw_rounding.c:120-168, 561-597)
    xs = [np.zeros(b.n) for b in prob.blocks]
    for j, c in enumerate(master.cols):
        if c["name"].startswith("lambda_") and sol["x"][j] != 0.0:
            _, ks, its = c["name"].split("_")
            xs[int(ks)] += sol["x"][j] * X[(int(ks), int(its))]
    return dict(objective=sol["obj"] + prob.shift, status=status, iterations=it, x=xs, trajectory=traj,
                n_master_cols=len(master.cols))
