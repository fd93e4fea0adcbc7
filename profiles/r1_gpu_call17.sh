timeout 900 python -m pytest tests/test_host_library.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_host_library.py 2>&1 | tail -3
