"""Pin the oracle end to end: a full Dantzig-Wolfe run (tests/dw_loop.py, HiGHS master) with the
oracle as the subproblem side must reproduce every objective the reference's own test-suite pins
(tests/dw-tests.sh:62,82,102,135 literals; ex_relaxed_solution vectors for the five deterministic
examples), to 1e-7 relative (north_star) — and the 7-significant-digit literal formatting."""
import numpy as np
import pytest

from conftest import EXAMPLES, load_example
from dw_loop import run_dw
from oracle import oracle as O


class OracleEngine:
    def __init__(self, prob, nthreads=4):
        self.b = O.OracleBlocks(prob)
        self.nt = nthreads

    def initial(self):
        return self.b.initial(self.nt)

    def iterate(self, pi, phase_one):
        return self.b.iterate(pi, phase_one, self.nt)


@pytest.mark.parametrize("name", EXAMPLES)
def test_examples_reach_reference_objective(name, golden):
    prob = load_example(name)
    res = run_dw(prob, OracleEngine(prob))
    exp = golden["objective"][name]
    assert res["status"] == "optimal"
    assert abs(res["objective"] - exp) <= 1e-7 * (1 + abs(exp))
    assert abs(res["objective"] - golden["highs"][name]) <= 1e-7 * (1 + abs(exp))
    if name in golden["objective_literal"]:
        assert "%e" % res["objective"] == golden["objective_literal"][name]
    if name in golden["relaxed_solution"]:
        sol = golden["relaxed_solution"][name]
        for k, b in enumerate(prob.blocks):
            for j, nm in enumerate(b.col_names):
                assert "%f" % res["x"][k][j] == "%f" % sol[nm], (nm, res["x"][k][j], sol[nm])


@pytest.mark.parametrize("gub", [True, False])
@pytest.mark.parametrize("name", ["book_dantzig", "four_sea", "web_trick", "book_lasdon", "single_sub"])
def test_product_master_reaches_reference_objective(name, gub, golden):
    """the whole DW loop on the CPU: oracle subproblems + the PRODUCT's master LP (GUB working basis / dense inverse)"""
    from dw_loop import ProductMaster, run_dw
    prob = load_example(name)
    res = run_dw(prob, OracleEngine(prob), master=ProductMaster(prob, gub=gub))
    exp = golden["objective"][name]
    assert res["status"] == "optimal"
    assert abs(res["objective"] - exp) <= 1e-7 * (1 + abs(exp))
