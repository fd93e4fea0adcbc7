timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py > gpurun_out/r2c_bench_B.json 2> gpurun_out/r2c_B.err; echo "B rc=$?"
python bench.py --config A > gpurun_out/r2c_bench_A.json 2> gpurun_out/r2c_A.err; echo "A rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python - <<PY
import json
for t in ("B","A"):
    d=json.loads(open("gpurun_out/r2c_bench_%s.json"%t).read().strip().splitlines()[-1])
    print(t, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["dw_time_to_optimal"], d["clocks"])
PY
