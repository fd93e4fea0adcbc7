for v in main w3 w5; do
lib=$PWD/dwsolver_b200/lib/libdwgpu.so; [ $v != main ] && lib=$PWD/alt/libdwgpu_$v.so
DWGPU_LIB=$lib DWGPU_TG_THREADS=128 python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1y_bench_B_$v.json 2> gpurun_out/r1y.err; echo "$v rc=$?"
done
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1y_bench_B_t64.json 2>> gpurun_out/r1y.err
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r1y_bench_B_*.json
