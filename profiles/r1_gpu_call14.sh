echo "== dual K=2 1536 G2"; DWGPU_DEBUG=1 timeout 300 python profiles/san_c.py 2 1536 3072 2 50 2>&1 | grep -E "debug|status|Error" | tail -8
echo "== dual K=4 2048 G4"; DWGPU_DEBUG=1 timeout 300 python profiles/san_c.py 4 2048 4096 4 50 2>&1 | grep -E "debug|status|Error" | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or path" 2>&1 | tail -3
