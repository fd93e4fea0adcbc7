"""dwsolver_b200 — ctypes mirrors of the B200 engine (`engine`, `master`) and of the drop-in host library (`solver`),
the block-angular problem model and the synthetic generators of BASELINE.json's configurations (`problem`).
The product is the native code under csrc/ and host/; these modules are what the tests and bench.py drive it with."""
