set -x
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_live_engines or mixed_shapes or edge_cases" 2>&1 | tail -15
for i in 1 2 3; do timeout 60 python -m pytest tests/test_host_library.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -3; done
