python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1p_bench_B_main.json 2> gpurun_out/r1p.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_occ7.so python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1p_bench_B_occ7.json 2>> gpurun_out/r1p.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_r144.so python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1p_bench_B_r144.json 2>> gpurun_out/r1p.err; echo "rc=$?"
grep -H -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1p_bench_*.json
