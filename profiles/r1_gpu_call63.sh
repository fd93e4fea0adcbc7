set -x
timeout 200 python -m pytest tests/test_host_library.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4
