"""Per-op geometry sweep on device-resident data: python tools/sweep_ops.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cvortex_b200 import api  # noqa: E402
from util import op_cases, vort_cases  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
api.initialise()
be = api.backend()
rng = np.random.default_rng(1)
peak = be.sm_count(0) * 128 * 1.965e9
st = torch.cuda.current_stream().cuda_stream
for op, reg in op_cases() + vort_cases():
    info = be.op_info(op, reg)
    src = rng.uniform(0, 10, (n, info["src_cols"])).astype(np.float32)
    src[:, -1] = 0.01
    if op.startswith("F3D"):
        src[:, 3:6] = src[:, 0:3] + rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
        src[:, 6] = rng.uniform(0, 10, n)
    tgt = src if info["tgt_cols"] == info["src_cols"] and not op.startswith("F3D") else rng.uniform(0, 10, (n, info["tgt_cols"])).astype(np.float32)
    s, t = torch.from_numpy(src).cuda(), torch.from_numpy(np.ascontiguousarray(tgt)).cuda()
    out = torch.empty((n, info["out_cols"]), device="cuda")
    line = f"{op:20s} {reg:12s} L={info['lane_ops']:2d} S={info['sfu_ops']}"
    for T in (4, 8, 2):
        be.tune(T, 0)
        best = 1e9
        for _ in range(3):
            be.m2m(op, reg, 0, st, s, n, t, n, out, 0.02, 1.0)
            torch.cuda.synchronize()
            best = min(best, be.last_pair_kernel_ms(0))
        rate = n * n / best / 1e6
        line += f" | T={T}: {rate:7.1f} Gpair/s {100 * rate * 1e9 * info['lane_ops'] / peak:5.1f}% fp32 {100 * rate * 1e9 * info['sfu_ops'] / (peak / 8):5.1f}% sfu"
    be.tune(0, 0)
    print(line, flush=True)
