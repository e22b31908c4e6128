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
