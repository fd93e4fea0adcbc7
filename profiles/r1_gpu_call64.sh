set -x
timeout 100 python -m pytest tests/test_host_library.py -m gpu -x -q -k "fewer_blocks" 2>&1 | tail -40
mkdir -p /tmp/bd && python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import load_example
from dwsolver_b200 import problem as P
P.write_block_angular(load_example('book_dantzig'), '/tmp/bd')"
cd /tmp/bd
for d in 0 0,0 0,0,0; do echo "== devices $d"; DWSOLVER_DEVICES=$d timeout 60 /root/repo/dwsolver_b200/bin/dwsolver -g guidefile 2>&1 | tail -6; done
