"""Few-target calls: pair-kernel (+ ordered finish) time against the number of equal runs.  python tools/few_sweep.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

api.initialise()
be = api.backend()
rng = np.random.default_rng(1)
st = torch.cuda.current_stream().cuda_stream
n = 1_000_000
src = torch.from_numpy(rng.uniform(0, 10, (n, 7)).astype(np.float32)).cuda()
sms = be.sm_count(0)
for m in (1, 16, 64, 255, 1024):
    tgt = torch.from_numpy(rng.uniform(0, 10, (m, 3)).astype(np.float32)).cuda()
    out = torch.empty((m, 3), device="cuda")
    line = f"M={m:5d}:"
    for T, grids in ((1, (296, 592, 1184, 2368, 3907)), (2, (148, 444, 888)), (8, (148, 296, 888))):
        for g in grids:
            be.tune(T, g)
            best = 1e9
            for _ in range(5):
                be.m2m("P3D_M2M_vel", "winckelmans", 0, st, src, n, tgt, m, out, 0.02)
                torch.cuda.synchronize()
                best = min(best, be.last_pair_kernel_ms(0))
            line += f" T{T}/{g}: {best * 1e3:6.0f}us"
    be.tune(0, 0)
    best = 1e9
    for _ in range(5):
        be.m2m("P3D_M2M_vel", "winckelmans", 0, st, src, n, tgt, m, out, 0.02)
        torch.cuda.synchronize()
        best = min(best, be.last_pair_kernel_ms(0))
    print(line + f" | planner {best * 1e3:6.0f}us {be.plan('P3D_M2M_vel', 0, n, m)}", flush=True)
