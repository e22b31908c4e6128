"""Few-target regime: python tools/small_targets.py"""
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
for m in (1, 16, 64, 255, 1024, 4096, 16384, 65536):
    tgt = torch.from_numpy(rng.uniform(0, 10, (m, 3)).astype(np.float32)).cuda()
    out = torch.empty((m, 3), device="cuda")
    best = 1e9
    for _ in range(5):
        be.m2m("P3D_M2M_vel", "winckelmans", 0, st, src, n, tgt, m, out, 0.02)
        torch.cuda.synchronize()
        best = min(best, be.last_pair_kernel_ms(0))
    print(f"N=1M sources x M={m:6d} targets: pair kernel {best:8.3f} ms  {n * m / best / 1e6:8.1f} Gpair/s  plan {be.plan('P3D_M2M_vel', 0, n, m)}", flush=True)
