for lib in libdwgpu_nt128_occ4.so libdwgpu_nt128_occ3.so libdwgpu_nt96_occ4.so; do
  echo "== $lib"; DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 592 8192 2>&1 | grep "^K="
done
