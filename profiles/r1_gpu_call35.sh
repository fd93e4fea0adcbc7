timeout 900 python -m pytest tests/test_host_library.py tests/test_gpu_parity.py -m gpu -x -q -k "host or cli or dw_solve or history" 2>&1 | tail -4
timeout 600 python profiles/dw_time_to_optimal.py A --arrays --no-files 2>&1 | tail -1
