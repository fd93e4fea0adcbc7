timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench_B.json 2> gpurun_out/r1e_bench_B.err; echo rc=$?; tail -n 2 gpurun_out/r1e_bench_B.err
