timeout 1500 python profiles/dw_harness_time.py B gpu 2>&1 | tail -1
timeout 1500 python profiles/dw_harness_time.py B cpu 2>&1 | tail -1
