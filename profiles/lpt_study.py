#!/usr/bin/env python
"""List-scheduling study: given the per-block iteration counts of every round (profiles/dump_iterations.py), how
long is a launch under the hardware's greedy CTA dispatch for different block orders?  Block time model:
t = t0 + it * t1 (prologue/epilogue + pivots).  Orders: natural, LPT by the previous round's iterations (what the
engine does), LPT by an exponential average, and the clairvoyant LPT (lower bound of any predictor)."""
import heapq
import sys

import numpy as np

it = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/iters_B.npy").astype(np.float64)
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 7 * 148
t0, t1 = 16.0, 7.5        # us per block, us per pivot (profiles/r1m_phase_times_B.txt)


def makespan(times, order):
    h = [0.0] * slots
    heapq.heapify(h)
    end = 0.0
    for k in order:
        s = heapq.heappop(h)
        e = s + times[k]
        end = max(end, e)
        heapq.heappush(h, e)
    return end


R, K = it.shape
tot = {"natural": 0, "lpt_prev": 0, "lpt_ema": 0, "lpt_max2": 0, "clairvoyant": 0, "ideal": 0}
ema = np.zeros(K)
prev = np.zeros(K)
prev2 = np.zeros(K)
for r in range(R):
    times = t0 + it[r] * t1
    nat = np.arange(K)
    res = {
        "natural": makespan(times, nat),
        "lpt_prev": makespan(times, np.argsort(-prev, kind="stable")),
        "lpt_ema": makespan(times, np.argsort(-ema, kind="stable")),
        "lpt_max2": makespan(times, np.argsort(-np.maximum(prev, prev2), kind="stable")),
        "clairvoyant": makespan(times, np.argsort(-times, kind="stable")),
        "ideal": max(times.sum() / slots, times.max()),
    }
    for k in tot:
        tot[k] += res[k]
    if r < 8:
        print("round %2d  mean it %6.1f max %4d  " % (r - 1, it[r].mean(), it[r].max()) + "  ".join("%s %7.0f" % (k, v) for k, v in res.items()))
    prev2 = prev
    prev = it[r]
    ema = 0.5 * ema + it[r]
print("sum over the step (us): " + "  ".join("%s %.0f" % (k, v) for k, v in tot.items()))
