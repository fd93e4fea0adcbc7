python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-time-to-optimal > gpurun_out/plain_bench_u.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches_B.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-time-to-optimal > gpurun_out/ncu_bench_u.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2f_launches_B.csv
