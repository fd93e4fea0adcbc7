timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1f_bench_B_n2.json 2> gpurun_out/r1f_bench_B_n2.err; echo rc=$?; tail -n 3 gpurun_out/r1f_bench_B_n2.err | cut -c1-300
