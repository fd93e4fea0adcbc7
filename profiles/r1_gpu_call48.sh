timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py > gpurun_out/r1t_bench_B.json 2> gpurun_out/r1t_B.err; echo "B rc=$?"
python bench.py --config A > gpurun_out/r1t_bench_A.json 2> gpurun_out/r1t_A.err; echo "A rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
cat gpurun_out/r1t_bench_B.json | cut -c1-1500
