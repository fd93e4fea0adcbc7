timeout 600 python profiles/dw_time_to_optimal.py A --arrays 2>&1 | tail -1
timeout 1500 python profiles/dw_time_to_optimal.py B --arrays --no-files 2>&1 | tail -1
