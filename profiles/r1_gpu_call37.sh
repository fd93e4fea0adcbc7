for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r1j_bench_B_n$n.json 2> gpurun_out/r1j_bench_B_n$n.err; echo "n=$n rc=$?"
done
