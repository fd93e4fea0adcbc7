set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or path" 2>&1 | tail -5
timeout 500 python profiles/run_c.py 64 2 2
