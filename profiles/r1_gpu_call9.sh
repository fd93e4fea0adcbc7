for lib in libdwgpu_occ6.so libdwgpu_occ5.so libdwgpu_occ4.so libdwgpu_occ6_s8.so libdwgpu_occ3_s8.so; do
  echo "== $lib"; DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 8192 2>&1 | grep "^K="
done
