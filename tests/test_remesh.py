"""Redistribution onto a grid and Pedrizzetti relaxation -- the steps either side of the
all-pairs sums.  CPU side: the oracle port against the reference's own
implementation (golden fixtures generated from it, and live when oracle/_ref is present), and
the product's host path (no accelerator enabled, or a user-defined interpolant) against both.
The GPU path is covered by tests/test_gpu_remesh.py."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from cvortex_b200.abi import RedistFunc
from util import REDISTS, assert_same_remesh, remesh_cases, remesh_particles

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_remesh.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def tol_for(cap):
    """Strength tolerance relative to the largest strength.  When the caller's array is too
    small nearly all the vorticity is dropped and handed back as one FP32 (reference) or FP64
    (here) running sum over thousands of nodes, divided among the few survivors: the sums agree
    to FP32 accumulation error, not to an ulp."""
    return 2e-6 if cap is None else 2e-5


def case_id(c):
    return f"{c[0]}d-{c[1]}-negl{c[2 + 1]}-cap{c[4]}"


# ---- the interpolants ---------------------------------------------------------------------

def test_interpolants_match_the_reference(oracle, product, ref):
    """Every library's cvtx_RedistFunc_* evaluates the same function, bit for bit, and carries
    the same radius (reference src/RedistFunc.cpp:36-96)."""
    us = np.concatenate([np.linspace(0, 2.5, 501), [0.5, 1.0, 1.5, 2.0, np.nextafter(np.float32(1), np.float32(0))]]).astype(np.float32)
    for name in REDISTS:
        pf = product.redistfunc(name)
        assert pf.radius == oracle.redist_radius(name)
        mine = np.array([pf.func(float(u)) for u in us], np.float32)
        want = np.array([oracle.redist(name, float(u)) for u in us], np.float32)
        assert np.array_equal(mine, want), name
        if ref is not None:
            rf = ref.redistfunc(name)
            assert rf.radius == pf.radius
            assert np.array_equal(np.array([rf.func(float(u)) for u in us], np.float32), want), name


@pytest.mark.parametrize("name", [n for n in REDISTS if n != "lambda0"])
def test_interpolants_partition_unity(oracle, name):
    """Sum over the grid of W(|x - k|) = 1 for any x: what makes redistribution conserve
    total vorticity.  (Lambda_0 loses a particle sitting exactly half way, as the reference.)"""
    for x in np.random.default_rng(3).uniform(0, 1, 200):
        total = sum(oracle.redist(name, abs(float(np.float32(x - k)))) for k in range(-3, 5))
        assert abs(total - 1.0) < 1e-6, (name, x, total)


# ---- oracle port == reference ---------------------------------------------------------------

@pytest.mark.parametrize("case", remesh_cases(), ids=case_id)
def test_oracle_matches_reference_golden(oracle, golden, case):
    dim, name, h, negl, cap = case
    key = f"remesh|{dim}|{name}|{negl}|{cap}"
    p, want = golden[key + "|in"], golden[key + "|out"]
    got = oracle.redistribute(p, name, h, negl, max_output=cap)
    assert_same_remesh(got, want, tol=tol_for(cap), what=key)
    assert abs(oracle.redistribute(p, name, h, negl, count_only=True) - int(golden[key + "|count"][0])) <= 2


@pytest.mark.parametrize("case", remesh_cases(), ids=case_id)
def test_host_path_matches_reference_golden(product, golden, case):
    """The product with no accelerator switched on runs its host stage (same arithmetic as the
    CUDA kernels, FP64 node sums)."""
    dim, name, h, negl, cap = case
    key = f"remesh|{dim}|{name}|{negl}|{cap}"
    p, want = golden[key + "|in"], golden[key + "|out"]
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    try:
        fn = product.P3D_redistribute_on_grid if dim == 3 else product.P2D_redistribute_on_grid
        got = fn(p, name, h, negl, max_output=cap)
        count = fn(p, name, h, negl, count_only=True)
    finally:
        for k in enabled:
            product.accelerator_enable(k)
    assert_same_remesh(got, want, tol=tol_for(cap), what=key)
    assert abs(count - int(golden[key + "|count"][0])) <= 2


def test_live_reference(oracle, product, ref):
    """Fresh inputs, larger than the fixtures, against the reference compiled here."""
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    try:
        for dim, h in ((3, 0.05), (2, 0.02)):
            for name in REDISTS:
                rng = np.random.default_rng(zlib.crc32(f"live{dim}{name}".encode()))
                p = remesh_particles(rng, 2500, dim)
                rfn = ref.P3D_redistribute_on_grid if dim == 3 else ref.P2D_redistribute_on_grid
                pfn = product.P3D_redistribute_on_grid if dim == 3 else product.P2D_redistribute_on_grid
                for negl, cap in ((0.0, None), (0.05, None), (0.01, 400)):
                    want = rfn(p, name, h, negl, max_output=cap)
                    assert_same_remesh(oracle.redistribute(p, name, h, negl, max_output=cap), want, tol=tol_for(cap),
                                       what=f"oracle {dim} {name} {negl} {cap}")
                    assert_same_remesh(pfn(p, name, h, negl, max_output=cap), want, tol=tol_for(cap),
                                       what=f"product {dim} {name} {negl} {cap}")
    finally:
        for k in enabled:
            product.accelerator_enable(k)


# ---- properties the domain offers ------------------------------------------------------------

@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("name", ["lambda1", "lambda2", "lambda3", "m4p"])
def test_total_vorticity_is_conserved(oracle, dim, name):
    """Interpolation conserves the total; pruning hands the dropped vorticity back evenly
    (reference src/P3D.cpp:651-663) -- so the sum survives any negligible_vort."""
    rng = np.random.default_rng(11)
    p = remesh_particles(rng, 2000, dim, signed=False)
    w = slice(3, 6) if dim == 3 else slice(2, 3)
    total_in = p[:, w].astype(np.float64).sum(0)
    for negl in (0.0, 0.3):
        out = oracle.redistribute(p, name, 0.07 if dim == 3 else 0.02, negl)
        assert np.allclose(out[:, w].astype(np.float64).sum(0), total_in, rtol=2e-5), (dim, name, negl)


@pytest.mark.parametrize("dim", [2, 3])
def test_linear_in_the_strengths(oracle, product, dim):
    """The grid depends on the positions only, every share is strength x weight: doubling the
    strengths (exact in FP32) doubles every created particle bit for bit, pruning included."""
    p = remesh_particles(np.random.default_rng(5), 700, dim)
    q = p.copy()
    w = slice(3, 6) if dim == 3 else slice(2, 3)
    q[:, w] *= 2
    for run in (lambda x: oracle.redistribute(x, "m4p", 0.09 if dim == 3 else 0.03, 0.05),
                lambda x: (product.P3D_redistribute_on_grid if dim == 3 else product.P2D_redistribute_on_grid)(
                    x, "m4p", 0.09 if dim == 3 else 0.03, 0.05)):
        a, b = run(p), run(q)
        assert np.array_equal(a[:, :dim], b[:, :dim]) and np.array_equal(2 * a[:, w], b[:, w])


def test_edge_cases(oracle, product):
    rf = product.redistfunc("m4p")
    # nothing in, nothing out
    empty = np.zeros((0, 7), np.float32)
    assert product.P3D_redistribute_on_grid(empty, "m4p", 0.1, count_only=True) == 0
    assert oracle.redistribute(empty, "m4p", 0.1, count_only=True) == 0
    # one particle: the grid is hung on the mean position, so it sits on a node (up to rounding)
    # and M4' leaves it there; the zero-weight part of the 5^3 stencil is not created
    one = np.array([[0.31, 0.62, 0.93, 1.0, -2.0, 0.5, 0.01]], np.float32)
    got = product.P3D_redistribute_on_grid(one, "m4p", 0.1)
    assert 1 <= len(got) <= 64 and len(got) == len(oracle.redistribute(one, "m4p", 0.1))
    assert np.array_equal(got[np.argmax(np.abs(got[:, 3]))][:3], one[0, :3])
    assert np.allclose(got[:, 3:6].sum(0), one[0, 3:6], rtol=1e-5)
    assert np.allclose(got[:, 6], np.float32(0.1) ** 3)
    # zero-strength particles create nothing
    zero = remesh_particles(np.random.default_rng(1), 50, 3)
    zero[:, 3:6] = 0
    assert product.P3D_redistribute_on_grid(zero, "lambda2", 0.1, count_only=True) == 0
    assert rf.radius == 2.0


def test_user_defined_interpolant_takes_the_host_path(product):
    """A caller-supplied cvtx_RedistFunc (here: a Python re-implementation of Lambda_1) cannot
    run on the GPU; the host stage calls it and must agree with the built-in."""
    calls = [0]

    def tent(u):
        calls[0] += 1
        return np.float32(1.0) - np.float32(u) if u <= 1.0 else 0.0

    cb = C.CFUNCTYPE(C.c_float, C.c_float)(tent)
    user = RedistFunc(func=cb, radius=1.0)
    p = remesh_particles(np.random.default_rng(9), 200, 3)
    want = product.P3D_redistribute_on_grid(p, "lambda1", 0.1, 0.0)
    got = product.P3D_redistribute_on_grid(p, user, 0.1, 0.0)
    assert calls[0] > 0
    assert_same_remesh(got, want, tol=0.0 if product.num_enabled_accelerators() == 0 else 2e-6)
    q = remesh_particles(np.random.default_rng(10), 200, 2)
    assert_same_remesh(product.P2D_redistribute_on_grid(q, user, 0.05, 0.0),
                       product.P2D_redistribute_on_grid(q, "lambda1", 0.05, 0.0), tol=2e-6)


# ---- relaxation ---------------------------------------------------------------------------

@pytest.mark.parametrize("reg", ["winckelmans", "gaussian", "planetary"])
def test_relaxation_matches_reference_golden(oracle, product, golden, reg):
    p, want = golden[f"relax|{reg}|in"], golden[f"relax|{reg}|out"]
    sigma, fdt = (float(v) for v in golden[f"relax|{reg}|par"])
    assert np.array_equal(oracle.pedrizzetti(p, fdt, reg, sigma), want), "oracle port is the reference arithmetic"
    got = product.P3D_pedrizzetti_relaxation(p, fdt, reg, sigma)
    assert np.array_equal(got[:, :3], p[:, :3]) and np.array_equal(got[:, 6], p[:, 6])
    scale = np.abs(want[:, 3:6]).max()
    assert np.abs(got[:, 3:6] - want[:, 3:6]).max() <= 1e-5 * scale


def test_relaxation_keeps_strength_and_turns_towards_the_field(oracle):
    """|alpha| changes by at most the blend, and fdt = 0 is the identity."""
    p = remesh_particles(np.random.default_rng(2), 300, 3)
    assert np.array_equal(oracle.pedrizzetti(p, 0.0, "gaussian", 0.2)[:, 3:6], p[:, 3:6])
    full = oracle.pedrizzetti(p, 1.0, "gaussian", 0.2)                 # fdt = 1: alpha -> |alpha| w/|w|
    assert np.allclose(np.linalg.norm(full[:, 3:6], axis=1), np.linalg.norm(p[:, 3:6], axis=1), rtol=1e-5)


def test_random_configurations_against_live_reference(oracle, product, ref):
    """150 seeded random calls -- 1 to 400 particles, every interpolant, 2-D and 3-D, boxes from
    1e-3 to 50 wide placed at 0 / -3 / 1000, grid spacings from 3 % to 5x the box, coincident
    particles, zero strengths, strengths scaled by 1e-20 .. 1e10, pruning on and off, output
    arrays that are too small -- against the reference compiled here.  Strengths are kept
    non-negative so that FP32 summation order (which differs in every implementation, the
    reference's own included) cannot cancel its way above the tolerance."""
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    rng = np.random.default_rng(2026)
    try:
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
            p[:, dim:dim + comps] = rng.uniform(0, 1, (n, comps)) * float(rng.choice([1.0, 1e-20, 1e10]))
            if rng.random() < 0.1:
                p[: n // 2, dim:dim + comps] = 0
            cap = None if rng.random() < 0.6 else int(rng.integers(1, 200))
            what = f"trial {trial}: {dim}D n={n} {name} box={box} h={h} negl={negl} cap={cap}"
            want = (ref.P3D_redistribute_on_grid if dim == 3 else ref.P2D_redistribute_on_grid)(p, name, h, negl, max_output=cap)
            got = (product.P3D_redistribute_on_grid if dim == 3 else product.P2D_redistribute_on_grid)(p, name, h, negl, max_output=cap)
            tol = 2e-6 if cap is None else 5e-5
            assert_same_remesh(got, want, tol=tol, what="product, " + what)
            assert_same_remesh(oracle.redistribute(p, name, h, negl, max_output=cap), want, tol=tol, what="oracle, " + what)
    finally:
        for k in enabled:
            product.accelerator_enable(k)
