"""In-library multi-GPU through the unchanged C ABI (ONE process): enable k accelerators with
cvtx_accelerator_enable and time cvtx_P3D_M2M_vel on host pointer arrays.
    python tools/multi_device_abi.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402
from cvortex_b200.abi import PointerRows  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
api.initialise()
lib, be = api.library(), api.backend()
rng = np.random.default_rng(20261017)
P = rng.uniform(0, 10, (n, 7)).astype(np.float32)
P[:, 6] = 0.01
X = rng.uniform(0, 10, (n, 3)).astype(np.float32)
src, out = PointerRows(P, 7), np.empty((n, 3), np.float32)
base = None
for k in (1, 2, 4, 8):
    if k > lib.num_accelerators():
        break
    for d in range(lib.num_accelerators()):
        (lib.accelerator_enable if d < k else lib.accelerator_disable)(d)
    lib.P3D_M2M_vel(src, X, "winckelmans", 0.02, out=out)            # warm-up (allocations)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        lib.P3D_M2M_vel(src, X, "winckelmans", 0.02, out=out)
        best = min(best, time.perf_counter() - t0)
    assert be.last_dispatch() == 1 and be.last_devices_used() == k
    ref = out.copy() if base is None else ref
    same = bool(np.array_equal(out, ref))
    base = best if base is None else base
    print(f"{k} accelerator(s) enabled: cvtx_P3D_M2M_vel {n} x {n} Winckelmans, host pointers in/out: {1e3 * best:8.2f} ms "
          f"{n * n / best / 1e9:9.1f} Gpair/s  speed-up {base / best:4.2f}x  bit-identical to 1 GPU: {same}", flush=True)
