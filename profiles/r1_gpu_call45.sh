python profiles/dump_iterations.py B 2>&1 | tail -n 2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1q_bench_B.json 2> gpurun_out/r1q.err; echo "rc=$?"
grep -H -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1q_bench_B.json
