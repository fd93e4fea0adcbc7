timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or path" 2>&1 | tail -3
timeout 600 python profiles/run_c.py 8 2 1
