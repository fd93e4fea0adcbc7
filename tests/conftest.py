import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


EXAMPLES = ["book_bertsimas", "book_bertsimas_double", "book_bertsimas_triple", "book_dantzig", "book_lasdon",
            "four_sea", "neg_y", "one_iter", "single_sub", "web_mitchell", "web_trick"]


def load_example(name):
    from dwsolver_b200 import problem as P
    return P.load_npz(os.path.join(GOLDEN_DIR, "examples", name + ".npz"))
