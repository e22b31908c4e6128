"""One few-target call (1M sources on M targets) for `ncu --metrics gpu__time_duration.sum`: shows the split
between pack_sources, the pair kernel and the reduction of the FP64 partials.  python tools/small_prof.py M"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from cvortex_b200 import api
api.initialise(); be = api.backend()
rng = np.random.default_rng(1); st = torch.cuda.current_stream().cuda_stream
n = 1_000_000; m = int(sys.argv[1])
src = torch.from_numpy(rng.uniform(0, 10, (n, 7)).astype(np.float32)).cuda()
tgt = torch.from_numpy(rng.uniform(0, 10, (m, 3)).astype(np.float32)).cuda()
out = torch.empty((m, 3), device="cuda")
for _ in range(3):
    be.m2m("P3D_M2M_vel", "winckelmans", 0, st, src, n, tgt, m, out, 0.02); torch.cuda.synchronize()
print(be.last_pair_kernel_ms(0), be.plan('P3D_M2M_vel', 0, n, m))
