set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in libdwgpu.so libdwgpu_occ6.so libdwgpu_s2r2_occ8.so libdwgpu_s2r4_occ7.so; do
  echo "== $lib"; DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 8192 2>&1 | grep "^K="
  echo "== $lib no reorder"; DWGPU_NO_REORDER=1 DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 8192 2>&1 | grep "^K="
done
