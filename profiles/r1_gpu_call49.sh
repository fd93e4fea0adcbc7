python -c "
import ctypes, os
print('hw', ctypes.CDLL('alt/libhc.so').hw_conc(), 'affinity', len(os.sched_getaffinity(0)))
import torch
print('after torch: hw', ctypes.CDLL('alt/libhc.so').hw_conc(), 'affinity', len(os.sched_getaffinity(0)))
"
DWSOLVER_MASTER_DEBUG=1 python profiles/dw_time_to_optimal.py B --arrays --no-files > gpurun_out/r1u_tto_B_default.json 2> gpurun_out/r1u_tto_B_default.err; echo "rc=$?"
tail -n 1 gpurun_out/r1u_tto_B_default.json | cut -c1-600
