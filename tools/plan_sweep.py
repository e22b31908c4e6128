"""Does the planner pick a good geometry for mid-size problems?  Sweeps (targets per thread, number of
equal runs = grid size) with cvtx_b200_tune and compares the best measured pair-kernel time with the
planner's own choice.   python tools/plan_sweep.py [n ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

api.initialise()
be = api.backend()
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)


def timed(op, reg, src, n, tgt, m, out, reps=7):
    best = 1e9
    for _ in range(reps):
        be.m2m(op, reg, 0, st, src, n, tgt, m, out, 0.02)
        torch.cuda.synchronize()
        best = min(best, be.last_pair_kernel_ms(0))
    return best


for n in [int(a) for a in sys.argv[1:]] or [3000, 10000, 30000, 80000]:
    for op, reg, tcols, ocols in (("P3D_M2M_vel", "winckelmans", 3, 3), ("P3D_M2M_dvort", "gaussian", 7, 3)):
        src = torch.from_numpy(rng.uniform(0, 10, (n, 7)).astype(np.float32)).cuda()
        tgt = src if tcols == 7 else torch.from_numpy(rng.uniform(0, 10, (n, 3)).astype(np.float32)).cuda()
        out = torch.empty((n, ocols), device="cuda")
        be.tune(0, 0)
        auto = timed(op, reg, src, n, tgt, n, out)
        plan = be.plan(op, 0, n, n)
        results = []
        sms = be.sm_count(0)
        for T in (8, 4, 2, 1):
            for chunks in [sms * k for k in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48)]:
                be.tune(T, chunks)
                results.append((timed(op, reg, src, n, tgt, n, out, reps=4), T, chunks))
        be.tune(0, 0)
        results.sort()
        best = results[0]
        print(f"{op}/{reg} n=m={n:6d}: planner {auto * 1e3:8.1f} us {plan} | best swept {best[0] * 1e3:8.1f} us (T={best[1]}, runs={best[2]}; next {results[1][0] * 1e3:.1f} us T={results[1][1]} runs={results[1][2]})"
              f" | planner / best = {auto / best[0]:.3f} | ideal at peak rate {n * n / (1574e9 if 'vel' in op else 800e9) * 1e6:7.1f} us", flush=True)
