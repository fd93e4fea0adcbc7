n=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r2b_bench_B_n$n.json 2> gpurun_out/r2b_bench_B_n$n.err; echo "n=$n rc=$?"
tail -n 3 gpurun_out/r2b_bench_B_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_bench_B_n$n.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["pivots"])
PY
