timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1k_bench_B_colmask.json 2> gpurun_out/r1k_colmask.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_nocm.so python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1k_bench_B_nocolmask.json 2> gpurun_out/r1k_nocm.err; echo "rc=$?"
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1k_bench_B_colmask2.json 2>> gpurun_out/r1k_colmask.err; echo "rc=$?"
python bench.py --config A --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1k_bench_A.json 2>> gpurun_out/r1k_colmask.err; echo "rc=$?"
grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1k_bench_*.json
