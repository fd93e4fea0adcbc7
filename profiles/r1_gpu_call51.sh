timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "path3_reads or global or config_b or mixed" 2>&1 | tail -n 5
DWGPU_POISON_TABLEAU=1 python bench.py --no-cpu-baseline --no-time-to-optimal > gpurun_out/r1w_bench_B_poison.json 2> gpurun_out/r1w.err; echo "rc=$?"
grep -H -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r1w_bench_B_poison.json
tail -n 3 gpurun_out/r1w.err
DWGPU_POISON_TABLEAU=1 timeout 600 python profiles/stress_dw.py 12 2>&1 | tail -n 2
