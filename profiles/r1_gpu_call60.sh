timeout 600 python -m pytest tests/test_host_library.py tests/test_oracle_dw.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python profiles/stress_dw.py 12 2>&1 | tail -n 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
