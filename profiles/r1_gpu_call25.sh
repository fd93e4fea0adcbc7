for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r1f_bench_B_n$n.json 2> gpurun_out/r1f_bench_B_n$n.err; echo "n=$n rc=$?"; tail -n 2 gpurun_out/r1f_bench_B_n$n.err | cut -c1-300
done
