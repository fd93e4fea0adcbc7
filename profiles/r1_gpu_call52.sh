for t in 128 256; do
DWGPU_TG_THREADS=$t python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1x_bench_B_t$t.json 2> gpurun_out/r1x.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r1x_bench_B_t$t.json").read().strip().splitlines()[-1])
print($t, d["ms_per_step"], [round(x,3) for x in d["roofline"]["kernel_ms_by_launch"]])
PY
done
