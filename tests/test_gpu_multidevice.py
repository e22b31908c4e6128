"""In-library multi-device sharding through the unchanged C ABI: enable several accelerators
(the reference's own switch, cvtx_accelerator_enable) and the targets of one call are split
over them.  Needs >= 2 GPUs; on a 1-GPU box only the single-device assertions run."""
import numpy as np
import pytest

from util import particles3d, points, rel_l2

pytestmark = pytest.mark.gpu


def test_targets_shard_over_enabled_accelerators(gpu, oracle):
    lib, dev = gpu
    n_acc = lib.num_accelerators()
    rng = np.random.default_rng(12)
    n, m = 60_000, 50_000                                  # 3e9 pairs: above the sharding threshold
    P, X = particles3d(rng, n, vol=0.01), points(rng, m, 3)
    one = lib.P3D_M2M_vel(P, X, "winckelmans", 0.02)
    assert dev.last_devices_used() == 1
    idx = np.arange(0, m, 97)
    assert rel_l2(one[idx], oracle.m2m("P3D_M2M_vel", P, np.ascontiguousarray(X[idx]), "winckelmans", 0.02)) <= 1e-5
    if n_acc < 2:
        pytest.skip("one GPU on this box: multi-device sharding not exercised")
    try:
        for k in range(n_acc):
            lib.accelerator_enable(k)
        assert lib.num_enabled_accelerators() == n_acc
        many = lib.P3D_M2M_vel(P, X, "winckelmans", 0.02)
        assert dev.last_devices_used() == n_acc
        # each target's sum is computed by exactly one device with the same kernel: identical bits
        assert np.array_equal(many, one)
        dv1 = lib.P3D_M2M_dvort(P, P[:m], "gaussian", 0.02)
        assert dev.last_devices_used() == n_acc
        small = lib.P3D_M2M_vel(P[:1000], X[:1000], "winckelmans", 0.02)     # too small to shard
        assert dev.last_devices_used() == 1
        assert rel_l2(small, oracle.m2m("P3D_M2M_vel", P[:1000], X[:1000], "winckelmans", 0.02)) <= 1e-5
    finally:
        for k in range(1, n_acc):
            lib.accelerator_disable(k)
    dv0 = lib.P3D_M2M_dvort(P, P[:m], "gaussian", 0.02)
    assert np.array_equal(dv0, dv1)


def test_sharded_sources_are_all_gathered_inside_the_library(gpu, oracle, torch_cuda):
    """cvtx_b200_m2m_sharded through the thin C ABI: device-resident source shards of UNEQUAL length, one
    per device, gathered over NCCL inside the library (BASELINE north_star), targets sharded too; every
    device must return the bits a single device computes from the concatenated sources.  With one GPU the
    same entry point runs with a single shard."""
    torch = torch_cuda
    lib, dev = gpu
    G = min(dev.device_count(), 8)
    rng = np.random.default_rng(21)
    n, m = 70_001, 9_000
    P = particles3d(rng, n, vol=0.01)
    T = np.ascontiguousarray(P[:m])
    for op, reg in (("P3D_M2M_dvort", "gaussian"), ("P3D_M2M_visc_dvort", "winckelmans")):
        src0, tgt0 = torch.from_numpy(P).to("cuda:0"), torch.from_numpy(T).to("cuda:0")
        want = torch.empty((m, 3), device="cuda:0")
        dev.m2m(op, reg, 0, torch.cuda.current_stream(0).cuda_stream, src0, n, tgt0, m, want, 0.02, 1.0)
        torch.cuda.synchronize(0)
        want = want.cpu().numpy()
        idx = np.arange(0, m, 45)
        assert rel_l2(want[idx], oracle.m2m(op, P, np.ascontiguousarray(T[idx]), reg, 0.02, 1.0)) <= 1e-5
        # unequal shards: the first device holds about half of the sources, the rest share the remainder
        cuts = [0] + [int(n * (0.5 + 0.5 * g / max(G - 1, 1))) for g in range(G - 1)] + [n] if G > 1 else [0, n]
        tcut = [m * g // G for g in range(G + 1)]
        shards = [torch.from_numpy(np.ascontiguousarray(P[cuts[g]:cuts[g + 1]])).to(f"cuda:{g}") for g in range(G)]
        tgts = [torch.from_numpy(np.ascontiguousarray(T[tcut[g]:tcut[g + 1]])).to(f"cuda:{g}") for g in range(G)]
        outs = [torch.empty((tcut[g + 1] - tcut[g], 3), device=f"cuda:{g}") for g in range(G)]
        for g in range(G):
            torch.cuda.synchronize(g)
        before = torch.cuda.current_device()
        dev.m2m_sharded(op, reg, list(range(G)), shards, [cuts[g + 1] - cuts[g] for g in range(G)],
                        tgts, [tcut[g + 1] - tcut[g] for g in range(G)], outs, 0.02, 1.0)
        assert torch.cuda.current_device() == before, "the library left the calling thread on another device"
        got = np.concatenate([o.cpu().numpy() for o in outs])
        assert np.array_equal(got, want), (op, reg, G)
    backend = dev.exchange_backend()
    print(f"{G} device(s); exchange backend: {backend}")
    if G > 1:
        assert "nccl" in backend or "peer" in backend


def test_the_enabled_set_can_change_between_calls(gpu, oracle):
    """cvtx_accelerator_enable / disable between calls ("re-create the communicator lazily"): every
    subset of devices returns the single-device bits, in any order of switching, and finalise / initialise in
    between starts over cleanly."""
    lib, dev = gpu
    n_acc = lib.num_accelerators()
    if n_acc < 2:
        pytest.skip("one GPU on this box")
    rng = np.random.default_rng(5)
    n, m = 90_000, 40_000
    P = particles3d(rng, n, vol=0.01)
    one = lib.P3D_M2M_visc_dvort(P, P[:m], "winckelmans", 0.02, 1.0)
    assert dev.last_devices_used() == 1
    try:
        for subset in ([0, 1], list(range(n_acc)), [n_acc - 1], [1, 0][:2], list(range(0, n_acc, 2)) or [0]):
            for k in range(n_acc):
                (lib.accelerator_enable if k in subset else lib.accelerator_disable)(k)
            got = lib.P3D_M2M_visc_dvort(P, P[:m], "winckelmans", 0.02, 1.0)
            assert dev.last_dispatch() == 1 and dev.last_devices_used() == len(set(subset)), (subset, dev.last_devices_used())
            assert np.array_equal(got, one), subset
        lib.finalise()
        lib.initialise()
        for k in range(n_acc):
            lib.accelerator_enable(k)
        got = lib.P3D_M2M_visc_dvort(P, P[:m], "winckelmans", 0.02, 1.0)
        assert np.array_equal(got, one)
    finally:
        for k in range(n_acc):
            (lib.accelerator_enable if k == 0 else lib.accelerator_disable)(k)
