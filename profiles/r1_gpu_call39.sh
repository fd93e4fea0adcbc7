timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1l_bench_B_fused.json 2> gpurun_out/r1l_fused.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_nofused.so python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1l_bench_B_nofused.json 2> gpurun_out/r1l_nofused.err; echo "rc=$?"
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1l_bench_B_fused2.json 2>> gpurun_out/r1l_fused.err; echo "rc=$?"
grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1l_bench_*.json
