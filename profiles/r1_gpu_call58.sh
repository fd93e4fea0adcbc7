python profiles/run_replay.py B 1 > gpurun_out/plain_replay_t.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 2 -c 1 -o gpurun_out/r2d_prof_B -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay_t.log 2>&1
tail -n 2 gpurun_out/ncu_replay_t.log
