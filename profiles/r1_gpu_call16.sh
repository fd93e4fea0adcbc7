set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1b_bench_B.json 2> gpurun_out/r1b_bench_B.err; echo rc=$?
timeout 300 python bench.py --config A > gpurun_out/r1b_bench_A.json 2> gpurun_out/r1b_bench_A.err; echo rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1b_launches_B.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
python profiles/run_replay.py B 1 > gpurun_out/plain_replay_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 2 -c 2 -o gpurun_out/r1b_prof_B -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay_b.log 2>&1
python profiles/san_c.py 8 1024 2048 16 50 > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_cluster -c 1 -o gpurun_out/r1b_prof_C -f python profiles/san_c.py 8 1024 2048 16 50 > gpurun_out/ncu_c.log 2>&1
tail -n 2 gpurun_out/ncu_c.log gpurun_out/plain_c.log
