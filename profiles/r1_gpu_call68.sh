timeout 33 python bench.py --config A --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_A.json 2> gpurun_out/r2h_bench_A.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2h_bench_A.json
