set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench_c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1c_launches_B.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c.log 2>&1
python profiles/run_replay.py B 1 > gpurun_out/plain_replay_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 2 -c 2 -o gpurun_out/r1c_prof_B -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay_c.log 2>&1
tail -n 2 gpurun_out/ncu_replay_c.log
