timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r2e_bench_B.json 2> gpurun_out/r2e.err; echo "rc=$?"
tail -n 2 gpurun_out/r2e.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench_B.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["pivots"], [round(x,3) for x in d["roofline"]["kernel_ms_by_launch"][:12]])
PY
