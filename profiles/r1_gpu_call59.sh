export DWGPU_LIB=$PWD/alt/libdwgpu_barrier.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 4
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r2e_bench_B.json 2> gpurun_out/r2e.err; echo "rc=$?"
unset DWGPU_LIB
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r2e_bench_B_head.json 2>> gpurun_out/r2e.err; echo "rc=$?"
tail -n 2 gpurun_out/r2e.err
python - <<PY
import json
for f in ("r2e_bench_B","r2e_bench_B_head"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["pivots"], [round(x,3) for x in d["roofline"]["kernel_ms_by_launch"][:8]])
PY
