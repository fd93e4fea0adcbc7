"""Pin the oracle's LP engine (oracle/orc_simplex.c — stands in for glp_simplex, whose source is
not in the reference tree) against scipy/HiGHS: random bounded LPs, every block of every
reference example under random prices, warm-start behaviour and degenerate/infeasible/
unbounded cases.  Tolerance: |z - z_highs| <= 1e-7 (1+|z|) (HiGHS itself stops at 1e-7);
integer-data cases are compared at 1e-9."""
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import EXAMPLES, load_example
from highs_util import solve_block, solve_rows
from oracle import oracle as O

OPT, NOFEAS, UNBND = 5, 4, 6


def rand_lp(rng, m, n, dens=0.4):
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < dens)
    x0 = rng.random(n)
    act = A @ x0
    kind = rng.integers(0, 3, m)
    rlb = np.where(kind == 0, act, np.where(kind == 1, -np.inf, act - rng.random(m)))
    rub = np.where(kind == 0, act, np.where(kind == 1, act + rng.random(m), np.inf))
    clb = np.zeros(n)
    cub = np.where(rng.random(n) < 0.7, 1.0 + rng.random(n), np.inf)
    free = rng.random(n) < 0.1
    clb[free] = -np.inf
    c = rng.standard_normal(n)
    c[np.isinf(cub) | free] = np.abs(c[np.isinf(cub) | free]) * 0  # keep it bounded
    return A, rlb, rub, clb, cub, c


@pytest.mark.parametrize("seed", range(12))
def test_random_lps_match_highs(seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(3, 25)), int(rng.integers(3, 40))
    A, rlb, rub, clb, cub, c = rand_lp(rng, m, n)
    S = sp.csr_matrix(A)
    lp = O.OracleLP(m, n, S.indptr, S.indices, S.data, rlb, rub, clb, cub)
    res = solve_rows(c, S, rlb, rub, clb, cub)
    st, z, x = lp.solve(c)
    assert res.status == 0 and st == OPT
    assert abs(z - res.fun) <= 1e-7 * (1 + abs(res.fun))
    assert lp.residual(x) <= 1e-9
    # warm re-solve with the same cost: zero pivots
    p0 = lp.counters()
    st2, z2, _ = lp.solve(c)
    assert st2 == OPT and z2 == z and lp.counters()[:2] == p0[:2]
    # objective-only change, warm start
    c2 = c + 0.5 * rng.standard_normal(n) * np.isfinite(cub) * np.isfinite(clb)
    res2 = solve_rows(c2, S, rlb, rub, clb, cub)
    st3, z3, x3 = lp.solve(c2)
    assert st3 == OPT and abs(z3 - res2.fun) <= 1e-7 * (1 + abs(res2.fun))
    assert lp.residual(x3) <= 1e-9


def test_infeasible_and_unbounded():
    # x1 + x2 <= 1, x1 + x2 >= 2
    lp = O.OracleLP(2, 2, [0, 2, 4], [0, 1, 0, 1], [1, 1, 1, 1.0], [-np.inf, 2], [1, np.inf], [0, 0], [np.inf, np.inf])
    st, _, _ = lp.solve(np.array([1.0, 1.0]))
    assert st == NOFEAS
    # min -x1, x1 - x2 <= 1, x >= 0 unbounded
    lp = O.OracleLP(1, 2, [0, 2], [0, 1], [1, -1.0], [-np.inf], [1], [0, 0], [np.inf, np.inf])
    st, _, _ = lp.solve(np.array([-1.0, 0.0]))
    assert st == UNBND


def test_empty_rows_and_fixed_columns():
    # no rows at all: optimum sits at the bounds
    lp = O.OracleLP(0, 3, [0], [], [], [], [], [0, -1, 2], [1, 1, 2])
    st, z, x = lp.solve(np.array([1.0, -1.0, 5.0]))
    assert st == OPT and np.allclose(x, [0, 1, 2]) and z == pytest.approx(9.0)


@pytest.mark.parametrize("name", EXAMPLES)
def test_example_blocks_under_random_prices(name):
    prob = load_example(name)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    blocks = O.OracleBlocks(prob)
    props = blocks.initial()
    for k, b in enumerate(prob.blocks):
        res = solve_block(b, b.c_file)
        assert res.status == 0 and props[k]["status"] == OPT
        assert abs(props[k]["obj"] - res.fun) <= 1e-9 * (1 + abs(res.fun))
    for trial in range(3):
        pi = rng.standard_normal(prob.L) * (trial + 1)
        phase_one = trial == 0
        props = blocks.iterate(pi, phase_one)
        for k, b in enumerate(prob.blocks):
            d = (0 if phase_one else b.c_master) - b.dense_d(prob.L).T @ pi
            res = solve_block(b, d)
            assert res.status == 0 and props[k]["status"] == OPT
            assert abs(props[k]["obj"] - res.fun) <= 1e-9 * (1 + abs(res.fun)), (name, k, trial)
            x = props[k]["x"]
            # proposal consistency: obj = d.x ; cost = c.x ; column = D x (exact nonzeros)
            assert abs(d @ x - props[k]["obj"]) <= 1e-9 * (1 + abs(props[k]["obj"]))
            assert abs(b.c_master @ x - props[k]["cost"]) <= 1e-12 * (1 + abs(props[k]["cost"]))
            y = np.zeros(prob.L)
            y[props[k]["ind"] - 1] = props[k]["val"]
            assert np.allclose(y, b.dense_d(prob.L) @ x, atol=1e-12)
            assert np.all(props[k]["val"] != 0.0)
