set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r1_bench_B.json 2> gpurun_out/r1_bench_B.err; echo rc=$?
python bench.py --config A > gpurun_out/r1_bench_A.json 2> gpurun_out/r1_bench_A.err; echo rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1_launches_B.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python profiles/run_replay.py B 1 > gpurun_out/plain_replay.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 2 -c 2 -o gpurun_out/r1_prof_B -f python profiles/run_replay.py B 1 > gpurun_out/ncu_replay.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 1 -c 2 -o gpurun_out/r1_prof_A -f python profiles/run_replay.py A 1 > gpurun_out/ncu_replayA.log 2>&1
ls -la gpurun_out
