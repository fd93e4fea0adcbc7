timeout 900 python -m pytest tests/test_host_library.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python profiles/dw_time_to_optimal.py A 2>&1 | tail -1
DWSOLVER_MASTER=dense timeout 600 python profiles/dw_time_to_optimal.py A 2>&1 | tail -1
