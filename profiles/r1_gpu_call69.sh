timeout 14 python profiles/tto_timing.py A 2>&1 | tail -8
