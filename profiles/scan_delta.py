#!/usr/bin/env python
"""Delta rounds against plain CTA rounds as a function of the number of blocks on the GPU (shard sizes of 8 / 4 / 2 / 1
GPUs at config B): replays the recorded trajectory on the first K blocks and prints the kernel time of the rounds from
10 on (few duals move) and of the whole replay.  Run once per setting of DWGPU_DELTA / DWGPU_DELTA_MIN_BLOCKS."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import load_trajectory  # noqa: E402
from dwsolver_b200 import problem as P  # noqa: E402
from dwsolver_b200.engine import Engine  # noqa: E402

pis, phases, _ = load_trajectory("B")
pis_dev = torch.from_numpy(pis).cuda()
for K in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]:
    prob = P.make_config("B", block_range=range(K))
    eng = Engine(prob, fetch_x=False)
    for rep in range(3):
        eng.reset()
        ms = []
        eng.initial_solve_device(); ms.append(max(eng.last_kernel_ms()))
        for r in range(len(phases)):
            eng.iterate_device(pis_dev[r].data_ptr(), phases[r]); ms.append(max(eng.last_kernel_ms()))
    late = ms[11:]
    print(f"DELTA={os.environ.get('DWGPU_DELTA', '1')} MIN={os.environ.get('DWGPU_DELTA_MIN_BLOCKS', 'default')} K={K:5d} "
          f"replay {sum(ms):7.3f} ms, rounds 10+: {sum(late):6.3f} ms, median late round {sorted(late)[len(late) // 2] * 1e3:6.1f} us", flush=True)
    eng.close()
