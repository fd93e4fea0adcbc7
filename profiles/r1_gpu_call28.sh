timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python profiles/scan_k.py B 1024 8192 2>&1 | grep "^K="
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1h_bench_B.json 2> gpurun_out/r1h_bench_B.err; echo rc=$?; tail -n 2 gpurun_out/r1h_bench_B.err
