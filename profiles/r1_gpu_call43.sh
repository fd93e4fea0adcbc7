for t in 1 2 4 8; do
DWSOLVER_MASTER_THREADS=$t DWSOLVER_MASTER_DEBUG=1 python profiles/dw_time_to_optimal.py B --arrays --no-files > gpurun_out/r1o_tto_B_t$t.json 2> gpurun_out/r1o_tto_B_t$t.err; echo "t=$t rc=$?"
tail -n 1 gpurun_out/r1o_tto_B_t$t.json | cut -c1-400
grep "seconds:" gpurun_out/r1o_tto_B_t$t.err | sed -n 2,3p
done
