for cfg in "2 512 1024 2" "2 1024 2048 2" "2 1536 3072 2" "2 2048 4096 2" "2 2048 4096 1"; do
  echo "== $cfg"; timeout 120 python profiles/san_c.py $cfg 1 2>&1 | tail -2
done
