"""Host-side logic of the multi-GPU layout (cvortex_b200/sharding.py): target partition and
the source all-gather, exercised with world_size-2 (and 3) `gloo` process groups on CPU.
The per-shard compute stand-in here is the oracle; what is tested is that sharding the
targets and gathering the sources reproduces the unsharded result bit for bit."""
import os
import socket

import numpy as np
import pytest

from cvortex_b200.sharding import all_ranges, target_range


def test_target_ranges_partition_exactly():
    for n in (0, 1, 7, 1000, 1_000_000, 4_000_003):
        for world in (1, 2, 3, 4, 8):
            r = all_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        target_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from cvortex_b200.sharding import allgather_rows, shard_rows, target_range
    from oracle.binding import Oracle
    from util import particles3d, points

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(99)                     # same particles on every rank
        P, X = particles3d(rng, n), points(rng, m, 3)
        mine = torch.from_numpy(np.ascontiguousarray(shard_rows(P, rank, world)))
        full = allgather_rows(mine, n).numpy()              # the exchange step
        assert np.array_equal(full, P), "all-gather must rebuild the source set in order"
        lo, hi = target_range(m, rank, world)
        ora = Oracle()
        part = ora.m2m("P3D_M2M_vel", full, np.ascontiguousarray(X[lo:hi]), "winckelmans", 0.3)
        np.save(os.path.join(tmp, f"part{rank}.npy"), part)
        dist.barrier()
        if rank == 0:
            whole = ora.m2m("P3D_M2M_vel", P, X, "winckelmans", 0.3)
            got = np.concatenate([np.load(os.path.join(tmp, f"part{r}.npy")) for r in range(world)])
            assert np.array_equal(got, whole), "sharded targets + replicated sources == unsharded"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,m", [(2, 1001, 333), (3, 1000, 301), (2, 64, 2)])
def test_sharded_equals_unsharded_over_gloo(tmp_path, world, n, m):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, m, str(tmp_path)), nprocs=world, join=True)
