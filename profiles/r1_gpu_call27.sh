set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r1g_bench_B.json 2> gpurun_out/r1g_bench_B.err; echo rc=$?
timeout 300 python bench.py --config A > gpurun_out/r1g_bench_A.json 2> gpurun_out/r1g_bench_A.err; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
