for v in 1 0; do
if [ $v = 1 ]; then export DWGPU_NO_REORDER=1; else unset DWGPU_NO_REORDER; fi
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r2a_bench_B_noreorder$v.json 2> gpurun_out/r2a.err; echo "rc=$?"
done
unset DWGPU_NO_REORDER
python bench.py --config A --no-cpu-baseline --no-time-to-optimal > gpurun_out/r2a_bench_A.json 2>> gpurun_out/r2a.err
python - <<PY
import json
for f in ("B_noreorder1","B_noreorder0","A"):
    d=json.loads(open("gpurun_out/r2a_bench_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], [round(x,3) for x in d["roofline"]["kernel_ms_by_launch"][:8]])
PY
