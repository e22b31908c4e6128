"""Input recipes and error metrics shared by the tests (and by bench.py)."""
from __future__ import annotations

import numpy as np

REGS = ("singular", "winckelmans", "planetary", "gaussian")
VISC_REGS = ("winckelmans", "gaussian")

# op -> (source cols, target cols, output cols, targets are particles?)
SHAPES = {
    "P3D_M2M_vel": (7, 3, 3, False), "P3D_M2M_dvort": (7, 7, 3, True), "P3D_M2M_visc_dvort": (7, 7, 3, True),
    "P3D_M2M_vort": (7, 3, 3, False), "P2D_M2M_vel": (4, 2, 2, False), "P2D_M2M_visc_dvort": (4, 4, 1, True),
    "F3D_M2M_vel": (7, 3, 3, False), "F3D_M2M_dvort": (7, 7, 3, True),
}


def rel_l2(a, b) -> float:
    """||a - b|| / ||b|| over the whole output array (the north-star metric)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    if den == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)


def upstream_per_target_ok(a, b, rel_acc=1e-5) -> bool:
    """The reference's own GPU-vs-CPU criterion, per target: |a-b|/|a+b| <= 1e-5 whenever
    |a+b| > 2e-35 (reference test/testsamecpugpuresultmany.h:39,99-106)."""
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    m = np.linalg.norm(a - b, axis=1)
    p = np.linalg.norm(a + b, axis=1)
    sel = p > 2e-35
    return bool(np.all(m[sel] / p[sel] <= rel_acc))


def particles3d(rng, n, box=10.0, vol=None):
    """cvtx_P3D rows as the reference's benchmark / tests build them: coords and vorticity
    uniform in [0, box), volume 0.01 (bench/bencharraysetup.c:43-58) or uniform [0, 0.01)
    (test/testsamecpugpuresultmany.h:69-78)."""
    p = rng.uniform(0.0, box, (n, 7)).astype(np.float32)
    p[:, 6] = rng.uniform(0.0, 0.01, n).astype(np.float32) if vol is None else np.float32(vol)
    return p


def particles2d(rng, n, box=10.0, area=None):
    p = rng.uniform(0.0, box, (n, 4)).astype(np.float32)
    p[:, 3] = rng.uniform(0.0, 0.01, n).astype(np.float32) if area is None else np.float32(area)
    return p


def filaments(rng, n, box=10.0, seg=None):
    """cvtx_F3D rows.  seg=None: both ends uniform in the box (the reference's test recipe,
    test/testsamecpugpuresultmany.h:80-88); else end = start + uniform(-seg, seg)^3."""
    f = rng.uniform(0.0, box, (n, 7)).astype(np.float32)
    if seg is not None:
        f[:, 3:6] = f[:, 0:3] + rng.uniform(-seg, seg, (n, 3)).astype(np.float32)
    return f


def points(rng, n, dim, box=10.0):
    return rng.uniform(0.0, box, (n, dim)).astype(np.float32)


def make_case(op, rng, n, m, box=10.0, self_targets=False):
    """(sources, targets) for `op`.  self_targets: targets are the first m sources (exercises
    the coincident-pair rule the way the reference's benchmark does for dvort / visc)."""
    sc, tc, _, tparticles = SHAPES[op]
    if op == "P3D_M2M_vort":
        op = "P3D_M2M_vel"
    if op.startswith("P2D"):
        src = particles2d(rng, n, box)
        tgt = src[:m].copy() if (tparticles and self_targets) else (particles2d(rng, m, box) if tparticles else points(rng, m, 2, box))
        if self_targets and not tparticles:
            tgt = np.ascontiguousarray(src[:m, :2])
    elif op.startswith("F3D"):
        src = filaments(rng, n, box, seg=0.1 * box / 10.0)
        tgt = particles3d(rng, m, box) if tparticles else points(rng, m, 3, box)
    else:
        src = particles3d(rng, n, box)
        tgt = src[:m].copy() if (tparticles and self_targets) else (particles3d(rng, m, box) if tparticles else points(rng, m, 3, box))
        if self_targets and not tparticles:
            tgt = np.ascontiguousarray(src[:m, :3])
    return src, tgt


def op_cases():
    """Every (op, regularisation) the reference accelerates."""
    out = []
    for reg in REGS:
        out += [("P3D_M2M_vel", reg), ("P3D_M2M_dvort", reg), ("P2D_M2M_vel", reg)]
    for reg in VISC_REGS:
        out += [("P3D_M2M_visc_dvort", reg), ("P2D_M2M_visc_dvort", reg)]
    out += [("F3D_M2M_vel", "singular"), ("F3D_M2M_dvort", "singular")]
    return out


def vort_cases():
    """cvtx_P3D_M2M_vort (singular has zeta = 0 everywhere)."""
    return [("P3D_M2M_vort", r) for r in ("winckelmans", "planetary", "gaussian")]


def call_abi(lib, op, src, tgt, reg, sigma, nu):
    """Run `op` through the cvtx_* C ABI of `lib` (a cvortex_b200.abi.CvtxLibrary)."""
    fn = getattr(lib, op)
    if op.startswith("F3D"):
        return fn(src, tgt)
    if op.endswith("visc_dvort"):
        return fn(src, tgt, reg, sigma, nu)
    return fn(src, tgt, reg, sigma)
