for cfg in "2 1536 3072 2"; do
  echo "== nodual $cfg"; NO_DUAL=1 timeout 120 python profiles/san_c.py $cfg 1 2>&1 | tail -2
  echo "== dual $cfg G4"; timeout 120 python profiles/san_c.py 2 1536 3072 4 1 2>&1 | tail -2
  echo "== dual 1 block $cfg"; timeout 120 python profiles/san_c.py 1 1536 3072 2 1 2>&1 | tail -2
done
