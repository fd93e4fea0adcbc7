timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "global or config_b or full_size or mixed or examples_fixed" 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1n_bench_B.json 2> gpurun_out/r1n.err; echo "rc=$?"
DWGPU_LIB=$PWD/alt/libdwgpu_timers.so python profiles/phase_times.py B > gpurun_out/r1n_phase_times_B.txt 2>> gpurun_out/r1n.err; echo "rc=$?"
grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1n_bench_B.json
tail -n 14 gpurun_out/r1n_phase_times_B.txt
