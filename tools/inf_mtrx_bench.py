"""Dense influence matrix throughput on device-resident data: python tools/inf_mtrx_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

api.initialise()
be = api.backend()
rng = np.random.default_rng(1)
for n, m in ((4096, 4096), (16384, 16384), (32768, 32768)):
    F = rng.uniform(0, 10, (n, 7)).astype(np.float32)
    F[:, 3:6] = F[:, 0:3] + rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
    f, x, d = torch.from_numpy(F).cuda(), torch.rand((m, 3), device="cuda") * 10, torch.rand((m, 3), device="cuda")
    out = torch.empty((m, n), device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(2):
        be.f3d_inf_mtrx(0, st.cuda_stream, f, n, x, d, m, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        be.f3d_inf_mtrx(0, st.cuda_stream, f, n, x, d, m, out)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    el = n * m / ms / 1e6
    print(f"inf_mtrx {n} filaments x {m} points: {ms:.3f} ms  {el:.1f} G elements/s  {el * 4 / 1e3:.2f} TB/s stored  "
          f"{100 * el * 1e9 * 41 / (148 * 128 * 1.965e9):.1f}% of FP32 peak (41 lane-ops/element)")
