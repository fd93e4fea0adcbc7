timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1v_bench_B.json 2> gpurun_out/r1v.err; echo "rc=$?"
grep -H -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1v_bench_B.json
tail -n 3 gpurun_out/r1v.err
