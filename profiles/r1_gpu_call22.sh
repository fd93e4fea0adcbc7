timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/scan_k.py B 1024 2048 4096 2>&1 | grep "^K="
