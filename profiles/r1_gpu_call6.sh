set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r6_bench_B.json 2> gpurun_out/r6_bench_B.err; echo rc=$?
DWGPU_LIB=$PWD/dwsolver_b200/lib/libdwgpu_occ6.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r6_bench_B_occ6.json 2> gpurun_out/r6_bench_B_occ6.err; echo rc=$?
tail -n 3 gpurun_out/r6_bench_B.err gpurun_out/r6_bench_B_occ6.err
