"""The device-resident restricted master (csrc/master_ipm.cu, include/dwgpu.h dwgpu_master_*) against HiGHS on random
DW-shaped masters: same optimum to 1e-7 relative, primal feasibility, dual feasibility of the returned row duals
(what the DW loop prices with).  Replaces the reference's glp_simplex(master_lp), src/dw_phases.c:219, 450."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from test_host_library import _gub_case

pytestmark = pytest.mark.gpu


def _matrix(cols, m):
    A = sp.lil_matrix((m, len(cols)))
    for j, (ix, vv) in enumerate(cols):
        for i, v in zip(ix, vv):
            A[i, j] = v
    return sp.csr_matrix(A)


@pytest.mark.parametrize("L,K,seed", [(8, 5, 1), (32, 60, 2), (144, 40, 3), (256, 300, 4), (100, 700, 5)])
def test_device_master_matches_highs(L, K, seed):
    from dwsolver_b200.master import DeviceMaster
    rng = np.random.default_rng(seed)
    cols, cost, rhs, _ = _gub_case(rng, L, K)
    A = _matrix(cols, L + K)
    ref = linprog(cost, A_eq=A, b_eq=rhs, bounds=[(0, None)] * len(cols), method="highs-ds")
    assert ref.status == 0
    m = DeviceMaster(L, K)
    m.set_rhs(rhs)
    m.add_cols(cols, cost)
    r = m.solve(tol=1e-9, max_iter=100)
    assert r["status"] == 0 and max(r["pinf"], r["dinf"], r["gap"]) <= 1e-7, r
    assert abs(r["obj"] - ref.fun) <= 1e-7 * (1 + abs(ref.fun)), (r["obj"], ref.fun)
    assert abs(r["dual_obj"] - ref.fun) <= 1e-7 * (1 + abs(ref.fun))
    x, y = r["x"], r["y"]
    assert x.min() >= 0.0
    assert np.abs(A @ x - rhs).max() <= 1e-7 * (1 + np.abs(rhs).max())
    assert (cost - A.T @ y).min() >= -1e-7 * (1 + np.abs(cost).max())      # reduced costs >= 0: y prices out every column
    # a second solve of the same model is bit-identical (fixed summation orders, no floating-point atomics)
    r2 = m.solve(tol=1e-9, max_iter=100)
    assert r2["iterations"] == r["iterations"] and np.array_equal(r2["x"], x) and np.array_equal(r2["y"], y)
    # columns deleted / costs changed: still the LP's optimum
    flags = np.zeros(len(cols), bool)
    flags[::7] = True
    flags[-2 * L:] = False                    # keep the slack pairs: the model stays feasible
    keep = ~flags
    m.del_cols(flags)
    cost2 = cost[keep] + rng.integers(0, 2, keep.sum())
    m.set_costs(0, cost2)
    A2 = A[:, keep]
    ref2 = linprog(cost2, A_eq=A2, b_eq=rhs, bounds=[(0, None)] * int(keep.sum()), method="highs-ds")
    if ref2.status == 0:
        r3 = m.solve(tol=1e-9, max_iter=100)
        assert r3["status"] == 0 and abs(r3["obj"] - ref2.fun) <= 1e-7 * (1 + abs(ref2.fun))
    m.close()
