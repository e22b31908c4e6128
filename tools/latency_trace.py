"""CVTX_B200_TRACE=1 python tools/latency_trace.py [n]: a few small calls with the library's own stage trace on stderr."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402
from cvortex_b200.abi import PointerRows  # noqa: E402

api.initialise()
be, lib = api.backend(), api.library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(1)
rows = rng.uniform(0, 1, (n, 7)).astype(np.float32)
rows[:, 6] = 0.01
mes = np.ascontiguousarray(rows[:, :3])
out = np.empty((n, 3), dtype=np.float32)
ptrs = PointerRows(rows, 7)
for _ in range(6):
    be.m2m_host("P3D_M2M_vel", "winckelmans", 0, rows, mes, 0.02, 0.0, out)
print("--- cvtx_P3D_M2M_vel (pointer arrays)", file=sys.stderr, flush=True)
for _ in range(4):
    lib.P3D_M2M_vel(ptrs, mes, "winckelmans", 0.02, out)
