set -x
timeout 300 python bench.py --config A > gpurun_out/r1c_bench_A.json 2> gpurun_out/r1c_bench_A.err; echo rc=$?
timeout 600 python bench.py > gpurun_out/r1c_bench_B.json 2> gpurun_out/r1c_bench_B.err; echo rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/r1c_ref_B.json 2> gpurun_out/r1c_ref_B.err; echo rc=$?
timeout 1500 python bench.py --config C --steps 1 --warmup 0 > gpurun_out/r1c_bench_C.json 2> gpurun_out/r1c_bench_C.err; echo rc=$?
tail -n 3 gpurun_out/r1c_bench_*.err gpurun_out/r1c_ref_B.err
