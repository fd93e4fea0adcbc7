set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r2_bench_B.json 2> gpurun_out/r2_bench_B.err; echo rc=$?
python bench.py --config A > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err; echo rc=$?
tail -3 gpurun_out/r2_bench_B.err gpurun_out/r2_bench_A.err
