set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r1i_bench_B.json 2> gpurun_out/r1i_bench_B.err; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
