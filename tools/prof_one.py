"""Launch one op a few times on device-resident data (for ncu):
    ncu --set full -k regex:m2m_kernel -c 1 -o gpurun_out/prof python tools/prof_one.py P3D_M2M_vel winckelmans 262144"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

op, reg, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
api.initialise()
be = api.backend()
info = be.op_info(op, reg)
rng = np.random.default_rng(1)
src = rng.uniform(0, 10, (n, info["src_cols"])).astype(np.float32)
src[:, -1] = 0.01
tgt = src if info["tgt_cols"] == info["src_cols"] else rng.uniform(0, 10, (n, info["tgt_cols"])).astype(np.float32)
s, t = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
out = torch.empty((n, info["out_cols"]), device="cuda")
for _ in range(reps):
    be.m2m(op, reg, 0, torch.cuda.current_stream().cuda_stream, s, n, t, n, out, 0.02, 1.0)
torch.cuda.synchronize()
ms = be.last_pair_kernel_ms(0)
print(f"{op}/{reg} n={n}: pair kernel {ms:.3f} ms, {n * n / ms / 1e6:.1f} Gpair/s, plan {be.plan(op, 0, n, n)}")
