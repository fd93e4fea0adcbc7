for thr in 128 64; do
  echo "== default build threads=$thr"; DWGPU_TG_THREADS=$thr python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-time-to-optimal 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['value']/1e6, d['pivots']['all_optimal'], [round(x,2) for x in d['roofline']['kernel_ms_by_launch'][:5]], d['roofline']['kernel_ms_by_launch'][20])"
  for occ in 8 9 10; do
    echo "== occ build $occ threads=$thr"; DWGPU_LIB=$PWD/alt/libdwgpu_occ$occ.so DWGPU_TG_THREADS=$thr python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-time-to-optimal 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['value']/1e6, d['pivots']['all_optimal'], [round(x,2) for x in d['roofline']['kernel_ms_by_launch'][:5]], d['roofline']['kernel_ms_by_launch'][20])"
  done
done
