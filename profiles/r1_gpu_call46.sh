timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1r_bench_B.json 2> gpurun_out/r1r.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_timers.so python profiles/phase_times.py B > gpurun_out/r1r_phase_times_B.txt 2>> gpurun_out/r1r.err; echo "rc=$?"
grep -H -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1r_bench_B.json
sed -n 30,36p gpurun_out/r1r_phase_times_B.txt
