set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-time-to-optimal > gpurun_out/plain_bench_s.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1s_launches_B.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-time-to-optimal > gpurun_out/ncu_bench_s.log 2>&1
python profiles/run_replay.py B 1 > gpurun_out/plain_replay_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 2 -c 2 -o gpurun_out/r1s_prof_B -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 25 -c 1 -o gpurun_out/r1s_prof_B_light -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay_s2.log 2>&1
tail -n 2 gpurun_out/ncu_replay_s.log gpurun_out/ncu_replay_s2.log
ls -la gpurun_out/*.ncu-rep
