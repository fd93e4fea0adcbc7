timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1z_bench_B.json 2> gpurun_out/r1z.err; echo "rc=$?"
tail -n 3 gpurun_out/r1z.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1z_bench_B.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
print(d["pivots"], d["gpu_launches"])
PY
