for lib in libdwgpu.so libdwgpu_nt32_occ9.so libdwgpu_nt32_occ9_s8.so libdwgpu_nt32_occ8_s6.so; do
  echo "== $lib"; DWGPU_LIB=$PWD/dwsolver_b200/lib/$lib timeout 300 python profiles/scan_k.py B 592 8192 2>&1 | grep "^K="
done
DWGPU_LIB=$PWD/dwsolver_b200/lib/libdwgpu_nt32_occ9.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config_b or dense_blocks_global" 2>&1 | tail -2
