echo "== dual K=2 1536 G2"; DWGPU_DEBUG=1 timeout 120 python profiles/san_c.py 2 1536 3072 2 1 2>&1 | grep -E "debug|status|Error" | tail -8
