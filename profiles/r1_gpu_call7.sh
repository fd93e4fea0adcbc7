set -x
for lib in libdwgpu.so libdwgpu_occ6.so libdwgpu_w0_occ8.so libdwgpu_w0_occ6.so; do
  echo "== $lib"; DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 592 1184 8192 2>&1 | grep "^K="
done
python profiles/scan_k.py B 1184 > gpurun_out/plain7.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_smem -s 7 -c 1 -o gpurun_out/r7_prof_B1184 -f python profiles/scan_k.py B 1184 > gpurun_out/ncu7.log 2>&1
tail -n 2 gpurun_out/ncu7.log
