"""Filament ops in both fast forms on device-resident data: python tools/f3d_modes_bench.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
api.initialise()
be = api.backend()
rng = np.random.default_rng(1)
st = torch.cuda.current_stream().cuda_stream
fil = rng.uniform(0, 10, (n, 7)).astype(np.float32)
fil[:, 3:6] = fil[:, 0:3] + rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
for op, tcols in (("F3D_M2M_vel", 3), ("F3D_M2M_dvort", 7)):
    tgt = rng.uniform(0, 10, (n, tcols)).astype(np.float32)
    s, t = torch.from_numpy(fil).cuda(), torch.from_numpy(tgt).cuda()
    out = torch.empty((n, 3), device="cuda")
    line = f"{op:14s} n = m = {n}:"
    for mode, name in ((0, "cancellation-free form (F3D_NEW)"), (1, "per-pair selecting form (F3D_WIDE)")):
        be.f3d_mode(mode)
        best = 1e9
        for _ in range(4):
            be.m2m(op, "singular", 0, st, s, n, t, n, out, 0.02)
            torch.cuda.synchronize()
            best = min(best, be.last_pair_kernel_ms(0))
        line += f"  {name}: {best:.2f} ms = {n * n / best / 1e6:.0f} Gpair/s;"
    be.f3d_mode(-1)
    print(line, flush=True)
