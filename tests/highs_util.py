"""Independent truth for the tests: scipy's bundled HiGHS."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def solve_rows(c, A, row_lb, row_ub, col_lb, col_ub, method="highs"):
    """min c'x s.t. row_lb <= A x <= row_ub, col bounds.  A: scipy sparse or dense."""
    A = sp.csr_matrix(A)
    eq = row_lb == row_ub
    up = np.isfinite(row_ub) & ~eq
    lo = np.isfinite(row_lb) & ~eq
    kw = {}
    if up.any() or lo.any():
        kw["A_ub"] = sp.vstack([A[up], -A[lo]], format="csr")
        kw["b_ub"] = np.concatenate([row_ub[up], -row_lb[lo]])
    if eq.any():
        kw["A_eq"] = A[eq]
        kw["b_eq"] = row_lb[eq]
    bounds = [(None if np.isinf(l) else l, None if np.isinf(u) else u) for l, u in zip(col_lb, col_ub)]
    res = linprog(c, bounds=bounds, method=method, **kw)
    return res


def solve_monolithic(prob):
    c, A, rlb, rub, clb, cub = prob.monolithic()
    res = solve_rows(c, A, rlb, rub, clb, cub)
    return dict(status=res.status, objective=(res.fun + prob.shift) if res.status == 0 else None, x=res.x)


def solve_block(block, cost):
    """Optimum of one block for a given priced cost vector (with the 3000 clamp applied)."""
    from dwsolver_b200.problem import effective_col_ub
    A = sp.csr_matrix((block.a_val, block.a_col, block.a_rowptr), shape=(block.m, block.n))
    res = solve_rows(cost, A, block.row_lb, block.row_ub, block.col_lb, effective_col_ub(block))
    return res
