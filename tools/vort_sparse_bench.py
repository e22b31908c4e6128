"""cvtx_P3D_M2M_vort on the particles a redistribution returns (Morton order): all tiles against marked tiles only.
    python tools/vort_sparse_bench.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvortex_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
api.initialise()
be = api.backend()
rng = np.random.default_rng(3)
P = rng.uniform(0, 10, (n, 7)).astype(np.float32)
P[:, 6] = 0.01
h = 10.0 * (1.0 / n) ** (1.0 / 3.0)                      # about one particle per cell
G = api.P3D_redistribute_on_grid(P, "m4p", h, 1e-3)
m = len(G)
sigma = 1.5 * h
print(f"{n} random particles -> {m} grid particles (h = {h:.4f}, Morton order); sigma = 1.5 h, cutoff cube {10 * sigma / 10:.3f} of a 10-wide box")
st = torch.cuda.current_stream().cuda_stream
src = torch.from_numpy(G).cuda()
tgt = src[:, :3].contiguous()
out = torch.empty((m, 3), device="cuda")
res = {}
for reg in ("gaussian", "winckelmans"):
    for sparse in (False, True):
        be.sparse_route(sparse)
        best = 1e9
        for _ in range(3):
            be.m2m("P3D_M2M_vort", reg, 0, st, src, m, tgt, m, out, sigma)
            torch.cuda.synchronize()
            best = min(best, be.last_pair_kernel_ms(0))
        res[(reg, sparse)] = (best, out.cpu().numpy().copy())
    same = np.array_equal(res[(reg, True)][1].view(np.uint32), res[(reg, False)][1].view(np.uint32))
    print(f"P3D_M2M_vort/{reg}: all tiles {res[(reg, False)][0]:8.2f} ms, marked tiles only {res[(reg, True)][0]:8.2f} ms "
          f"({res[(reg, False)][0] / res[(reg, True)][0]:.1f}x), same bits: {same}")
be.sparse_route(True)
import time  # noqa: E402
for reg in ("gaussian",):
    rows = src.clone()
    for sparse in (False, True):
        be.sparse_route(sparse)
        be.pedrizzetti_relaxation(reg, 0, st, rows, m, 0.1, sigma)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be.pedrizzetti_relaxation(reg, 0, st, rows, m, 0.1, sigma)      # (a null stream: the library's own; the sync below covers it)
        torch.cuda.synchronize()
        print(f"cvtx_b200_pedrizzetti_relaxation/{reg} on the same particles, sparse route {'on ' if sparse else 'off'}: {1e3 * (time.perf_counter() - t0):8.2f} ms")
be.sparse_route(True)
