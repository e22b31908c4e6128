"""GPU tests of redistribution onto a grid and relaxation, through the public cvtx_* ABI:
against the reference's own numbers (golden fixtures, and live when oracle/_ref travelled),
against the oracle port, and -- bit for bit -- against the library's own host stage."""
import os
import zlib

import numpy as np
import pytest

from test_remesh import tol_for
from util import REDISTS, assert_same_remesh, remesh_cases, remesh_particles

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_remesh.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def fn_of(lib, dim):
    return lib.P3D_redistribute_on_grid if dim == 3 else lib.P2D_redistribute_on_grid


def on_host(product, call):
    """Run `call` with every accelerator switched off (the library's host stage)."""
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    try:
        return call()
    finally:
        for k in enabled:
            product.accelerator_enable(k)


@pytest.mark.parametrize("case", remesh_cases(), ids=lambda c: f"{c[0]}d-{c[1]}-negl{c[3]}-cap{c[4]}")
def test_gpu_matches_reference_golden(gpu, golden, case):
    product, dev = gpu
    dim, name, h, negl, cap = case
    key = f"remesh|{dim}|{name}|{negl}|{cap}"
    p, want = golden[key + "|in"], golden[key + "|out"]
    before = dev.kernel_launches()
    got = fn_of(product, dim)(p, name, h, negl, max_output=cap)
    assert dev.last_dispatch() == 1 and dev.kernel_launches() - before >= 3, "the CUDA stage must have run"
    assert_same_remesh(got, want, tol=tol_for(cap), what=key)
    assert abs(fn_of(product, dim)(p, name, h, negl, count_only=True) - int(golden[key + "|count"][0])) <= 2


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("name", REDISTS)
def test_gpu_and_host_stage_give_the_same_bits(gpu, dim, name):
    """Same shares, same stable order, same FP64 node sums: the two stages are interchangeable."""
    product, dev = gpu
    rng = np.random.default_rng(zlib.crc32(f"bits{dim}{name}".encode()))
    n = 30011
    p = remesh_particles(rng, n, dim)
    h = float(np.cbrt(2.0 / n)) if dim == 3 else float(np.sqrt(2.0 / n))
    for negl, cap in ((0.0, None), (1e-4, None), (0.01, n // 3)):
        got = fn_of(product, dim)(p, name, h, negl, max_output=cap)
        assert dev.last_dispatch() == 1
        want = on_host(product, lambda: fn_of(product, dim)(p, name, h, negl, max_output=cap))
        assert dev.last_dispatch() == 0
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, name, negl, cap)
    # the same particles on a grid 12x finer per axis: too sparse for the dense route, so the device
    # sorts shares instead (32-bit codes); still the same bits as the host stage
    fine = h / 12
    got = fn_of(product, dim)(p, name, fine, 1e-4)
    want = on_host(product, lambda: fn_of(product, dim)(p, name, fine, 1e-4))
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, name, "sparse")


def test_gpu_matches_oracle_and_live_reference(gpu, oracle, ref):
    product, _ = gpu
    for dim, h in ((3, 0.05), (2, 0.02)):
        for name in REDISTS:
            rng = np.random.default_rng(zlib.crc32(f"gpu-live{dim}{name}".encode()))
            p = remesh_particles(rng, 4000, dim)
            for negl, cap in ((0.0, None), (0.05, None), (0.01, 500)):
                got = fn_of(product, dim)(p, name, h, negl, max_output=cap)
                assert_same_remesh(got, oracle.redistribute(p, name, h, negl, max_output=cap), tol=tol_for(cap),
                                   what=f"oracle {dim} {name} {negl} {cap}")
                if ref is not None:
                    assert_same_remesh(got, fn_of(ref, dim)(p, name, h, negl, max_output=cap), tol=tol_for(cap),
                                       what=f"reference {dim} {name} {negl} {cap}")


@pytest.mark.parametrize("dim", [2, 3])
def test_fine_grids_use_wide_codes(gpu, oracle, dim):
    """More than 2^10 (3-D) / 2^16 (2-D) nodes per axis: the device switches to 64-bit node
    codes.  Sparse particles on a very fine grid, so the node count stays small."""
    product, dev = gpu
    p = remesh_particles(np.random.default_rng(31), 3000, dim)
    h = 4e-4 if dim == 3 else 1e-5
    for name in ("lambda1", "m4p"):
        got = fn_of(product, dim)(p, name, h, 0.0)
        assert dev.last_dispatch() == 1
        want = on_host(product, lambda: fn_of(product, dim)(p, name, h, 0.0))
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert_same_remesh(got, oracle.redistribute(p, name, h, 0.0), what=f"fine {dim} {name}")


@pytest.mark.parametrize("dim,n", [(3, 1_000_000), (2, 1_000_000)])
def test_benchmark_size_conserves_vorticity_and_is_deterministic(gpu, dim, n):
    """The reference benchmark's largest case (bench/benchredistribution.c:49-61: a million
    particles in the unit box, about two per cell, M4', negligible_vort 1e-4, room for 4n)."""
    product, _ = gpu
    p = remesh_particles(np.random.default_rng(77), n, dim, signed=False)
    h = float(np.cbrt(2.0 / n))
    out = fn_of(product, dim)(p, "m4p", h, 1e-4, max_output=4 * n)
    w = slice(3, 6) if dim == 3 else slice(2, 3)
    assert 0 < len(out) <= 4 * n
    assert np.allclose(out[:, w].astype(np.float64).sum(0), p[:, w].astype(np.float64).sum(0), rtol=1e-5)
    lo, hi = p[:, :dim].min(0) - 2.5 * h, p[:, :dim].max(0) + 2.5 * h
    assert np.all(out[:, :dim] >= lo) and np.all(out[:, :dim] <= hi)
    again = fn_of(product, dim)(p, "m4p", h, 1e-4, max_output=4 * n)
    assert np.array_equal(out.view(np.uint32), again.view(np.uint32))


@pytest.mark.parametrize("reg", ["winckelmans", "gaussian", "planetary"])
def test_relaxation_on_the_gpu(gpu, golden, oracle, reg):
    product, dev = gpu
    p, want = golden[f"relax|{reg}|in"], golden[f"relax|{reg}|out"]
    sigma, fdt = (float(v) for v in golden[f"relax|{reg}|par"])
    got = product.P3D_pedrizzetti_relaxation(p, fdt, reg, sigma)
    assert dev.last_dispatch() == 1, "the vorticity field must have come from the CUDA kernel"
    assert np.array_equal(got[:, :3], p[:, :3]) and np.array_equal(got[:, 6], p[:, 6])
    assert np.abs(got[:, 3:6] - want[:, 3:6]).max() <= 1e-5 * np.abs(want[:, 3:6]).max()
    big = remesh_particles(np.random.default_rng(4), 20000, 3)
    got = product.P3D_pedrizzetti_relaxation(big, 0.2, reg, 0.05)
    want = oracle.pedrizzetti(big, 0.2, reg, 0.05)
    assert np.abs(got[:, 3:6] - want[:, 3:6]).max() <= 1e-5 * np.abs(want[:, 3:6]).max()


# ---- particles resident on the device (cvtx_b200_redistribute, include/cvtx_b200.h) ----------

@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("name", REDISTS)
def test_device_resident_matches_the_host_array_entry_point(gpu, torch_cuda, dim, name):
    """Same grid, same nodes, same order, same pruning sums in the same order: the same bits."""
    torch = torch_cuda
    product, dev = gpu
    rng = np.random.default_rng(zlib.crc32(f"resident{dim}{name}".encode()))
    n = 50021
    p = remesh_particles(rng, n, dim)
    h = float(np.cbrt(2.0 / n)) if dim == 3 else float(np.sqrt(2.0 / n))
    rows = torch.from_numpy(p).cuda()
    cols = p.shape[1]
    for negl, cap in ((0.0, 4 * n), (1e-3, 4 * n), (0.01, n // 4)):
        want = fn_of(product, dim)(p, name, h, negl, max_output=cap)
        out = torch.full((cap, cols), float("nan"), device="cuda")
        before = dev.kernel_launches()
        k = dev.redistribute(dim, name, 0, torch.cuda.current_stream().cuda_stream, rows, n, h, negl, out, cap)
        assert dev.kernel_launches() - before >= 6
        got = out[:k].cpu().numpy()
        assert torch.isnan(out[k:]).all(), "nothing may be written past the returned count"
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, name, negl, cap)
        # count-only mode: the reference's NULL idiom
        assert dev.redistribute(dim, name, 0, None, rows, n, h, negl) == fn_of(product, dim)(p, name, h, negl, count_only=True)


def test_device_resident_edge_cases(gpu, torch_cuda):
    torch = torch_cuda
    _, dev = gpu
    from cvortex_b200.device import BackendError
    rows = torch.zeros((10, 7), device="cuda")
    rows[:, :3] = torch.rand((10, 3), device="cuda")
    out = torch.zeros((100, 7), device="cuda")
    assert dev.redistribute(3, "m4p", 0, None, rows, 0, 0.1, 0.0, out, 100) == 0          # no particles
    assert dev.redistribute(3, "m4p", 0, None, rows, 10, 0.1, 0.0, out, 100) == 0         # no vorticity
    with pytest.raises(BackendError):
        dev.redistribute(3, "m4p", 0, None, rows, 10, 0.0, 0.0, out, 100)                 # spacing must be > 0
    with pytest.raises(BackendError):
        dev.redistribute(4, "m4p", 0, None, rows, 10, 0.1, 0.0, out, 100)
    rows[:, 3] = 1.0
    with pytest.raises(BackendError):
        dev.redistribute(3, "m4p", 0, None, rows, 10, 1e-8, 0.0, out, 100)                # 2^21 nodes per axis exceeded
    fine = dev.redistribute(3, "lambda1", 0, None, rows, 10, 4e-4, 0.0, out, 100)         # 64-bit codes
    assert 10 <= fine <= 80


def test_random_configurations_same_bits_as_the_host_stage(gpu, torch_cuda):
    """The configurations of tests/test_remesh.py's randomized test (1-400 particles, every
    interpolant, tiny to huge cells, coincident particles, zero and extreme strengths, pruning,
    small output arrays), signed strengths this time: whatever route and code width the device
    picks, the host-array call returns the host stage's bits, and the device-pointer call the
    same particles."""
    torch = torch_cuda
    product, dev = gpu
    rng = np.random.default_rng(7)
    for trial in range(150):
        dim = int(rng.choice([2, 3]))
        n = int(rng.choice([1, 2, 3, 5, 17, 100, 400]))
        name = str(rng.choice(REDISTS))
        comps = 3 if dim == 3 else 1
        box = float(rng.choice([1e-3, 1.0, 50.0]))
        h = box * float(rng.choice([0.03, 0.2, 1.0, 5.0]))
        negl = float(rng.choice([0.0, 1e-4, 0.3, 0.9]))
        p = np.zeros((n, dim + comps + 1), np.float32)
        p[:, :dim] = float(rng.choice([0.0, -3.0, 1000.0])) + rng.uniform(0, box, (n, dim))
        if rng.random() < 0.2:
            p[:, :dim] = p[0, :dim]
        p[:, dim:dim + comps] = rng.uniform(-1, 1, (n, comps)) * float(rng.choice([1.0, 1e-20, 1e10]))
        if rng.random() < 0.1:
            p[: n // 2, dim:dim + comps] = 0
        cap = None if rng.random() < 0.6 else int(rng.integers(1, 200))
        what = f"trial {trial}: {dim}D n={n} {name} box={box} h={h} negl={negl} cap={cap}"
        got = fn_of(product, dim)(p, name, h, negl, max_output=cap)
        want = on_host(product, lambda: fn_of(product, dim)(p, name, h, negl, max_output=cap))
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), what
        room = cap if cap is not None else max(len(want), 1)
        out = torch.full((room, p.shape[1]), float("nan"), device="cuda")
        k = dev.redistribute(dim, name, 0, None, torch.from_numpy(p).cuda(), n, h, negl, out, room)
        resident = out[:k].cpu().numpy()
        assert resident.shape == want.shape and np.array_equal(resident.view(np.uint32), want.view(np.uint32)), "device-resident, " + what


@pytest.mark.parametrize("reg", ["winckelmans", "gaussian", "planetary"])
def test_device_resident_relaxation(gpu, torch_cuda, oracle, reg):
    """cvtx_b200_pedrizzetti_relaxation: same kernel for the field, same FP32 blend as the
    host-array entry point -- the same bits; and within 1e-5 of the reference arithmetic."""
    torch = torch_cuda
    product, dev = gpu
    p = remesh_particles(np.random.default_rng(12), 20000, 3)
    want = product.P3D_pedrizzetti_relaxation(p, 0.3, reg, 0.05)
    rows = torch.from_numpy(p).cuda()
    dev.pedrizzetti_relaxation(reg, 0, torch.cuda.current_stream().cuda_stream, rows, len(p), 0.3, 0.05)
    torch.cuda.synchronize()
    got = rows.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ref = oracle.pedrizzetti(p, 0.3, reg, 0.05)
    assert np.abs(got[:, 3:6] - ref[:, 3:6]).max() <= 1e-5 * np.abs(ref[:, 3:6]).max()
    dev.pedrizzetti_relaxation(reg, 0, None, rows, 0, 0.3, 0.05)          # nothing to do
