set -x
timeout 170 python -m pytest tests/test_host_library.py -m gpu -x -q 2>&1 | tail -4
timeout 90 python profiles/input_path_timing.py A > gpurun_out/r2g_input_path_A.json 2> gpurun_out/r2g_input_path_A.err; echo "rc=$?"; cat gpurun_out/r2g_input_path_A.json
