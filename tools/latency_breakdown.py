"""Where a small call's time goes (10k x 10k through cvtx_P3D_M2M_vel):
    python tools/latency_breakdown.py [n ...]
Four nested timings of the same (op, n, m), wall clock on the host, median of many calls, no L2 flush:
  kernel      the pair kernel alone (CUDA events inside the library)
  device      cvtx_b200_m2m on device-resident rows + stream synchronise   (launch + kernel + wake-up)
  host rows   cvtx_b200_m2m_host on flat host arrays                        (+ staging copy, H2D, result copy)
  cvtx_*      cvtx_P3D_M2M_vel on arrays of pointers                        (+ the pointer gather)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402
from cvortex_b200.abi import PointerRows  # noqa: E402

api.initialise()
be = api.backend()
lib = api.library()
rng = np.random.default_rng(1)
st = torch.cuda.current_stream().cuda_stream
sizes = [int(a) for a in sys.argv[1:] if not a.startswith("T=")] or [3000, 10000, 30000]
for a in sys.argv[1:]:
    if a.startswith("T="):
        be.tune(int(a[2:]), 0)           # force the targets per thread (experiments)


def median_us(fn, reps):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e6 * ts[len(ts) // 2], 1e6 * ts[len(ts) // 10]


for n in sizes:
    rows = rng.uniform(0, 1, (n, 7)).astype(np.float32)
    rows[:, 6] = 0.01
    mes = np.ascontiguousarray(rows[:, :3])
    d_src = torch.from_numpy(rows).cuda()
    d_tgt = torch.from_numpy(mes).cuda()
    d_out = torch.empty((n, 3), device="cuda")
    out = np.empty((n, 3), dtype=np.float32)
    ptrs = PointerRows(rows, 7)
    reps = 300 if n <= 10000 else 60

    def device_call():
        be.m2m("P3D_M2M_vel", "winckelmans", 0, st, d_src, n, d_tgt, n, d_out, 0.02)
        torch.cuda.synchronize()

    kern = []
    for _ in range(20):
        device_call()
        kern.append(be.last_pair_kernel_ms(0) * 1e3)
    kern.sort()
    dev = median_us(device_call, reps)
    host = median_us(lambda: be.m2m_host("P3D_M2M_vel", "winckelmans", 0, rows, mes, 0.02, 0.0, out), reps)
    full = median_us(lambda: lib.P3D_M2M_vel(ptrs, mes, "winckelmans", 0.02, out), reps)
    print(f"n = m = {n:6d}: kernel {kern[len(kern) // 2]:7.1f} us | device call + sync {dev[0]:7.1f} (p10 {dev[1]:.1f}) | "
          f"flat host rows {host[0]:7.1f} (p10 {host[1]:.1f}) | cvtx_P3D_M2M_vel {full[0]:7.1f} (p10 {full[1]:.1f}) | "
          f"plan {be.plan('P3D_M2M_vel', 0, n, n)}", flush=True)
