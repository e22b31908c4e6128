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


def upstream_per_target_rejected(a, b, rel_acc=1e-5):
    """Mask of the targets the reference's criterion rejects (see upstream_per_target_ok)."""
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    m = np.linalg.norm(a - b, axis=1)
    p = np.linalg.norm(a + b, axis=1)
    return (p > 2e-35) & (m > rel_acc * p)


class UpstreamRand:
    """The reference test program's own generator (reference test/testmain.c:87-92): MSVC's LCG on a
    32-bit int seeded with 0xF0F0F0F0, bits 16..30, modulo RAND_MAX -- with the author's RAND_MAX of
    32767 (oracle/rand_max_msvc.h)."""

    def __init__(self, seed=0xF0F0F0F0):
        self.v = seed & 0xFFFFFFFF

    def mrand(self):
        self.v = (self.v * 214013 + 2531011) & 0xFFFFFFFF
        signed = self.v - (1 << 32) if self.v >= (1 << 31) else self.v
        return ((signed >> 16) & 0x7FFF) % 32767

    def ints(self, n):
        return np.array([self.mrand() for _ in range(n)], dtype=np.float32)


def upstream_many_inputs(gen, n=1000, max_float=10.0):
    """One repeat of the reference's "Many particles with vorticity" inputs, in its order and with its
    arithmetic (reference test/testsamecpugpuresultmany.h:68-88, :335-341): `(float)mrand() /
    (float)(RAND_MAX / max_float)` for coordinates and strengths, `(float)mrand() / (float)(RAND_MAX /
    0.01)` for volumes / areas; particles, then filaments with both ends anywhere in the box, then 2-D
    particles.  The measurement points of the recipe are the particle positions."""
    d = np.float32(32767.0) / np.float32(max_float)          # int / float: single precision
    dv = np.float32(32767.0 / 0.01)                           # int / double, then the cast
    P = gen.ints(n * 7).reshape(n, 7)
    P[:, :6] /= d
    P[:, 6] /= dv
    F = gen.ints(n * 7).reshape(n, 7) / d
    P2 = gen.ints(n * 4).reshape(n, 4)
    P2[:, :3] /= d
    P2[:, 3] /= dv
    return np.ascontiguousarray(P), np.ascontiguousarray(F.astype(np.float32)), np.ascontiguousarray(P2)


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


def nasty_case(op, rng, n, m, nonfinite=True):
    """Self-interaction plus everything that makes a guard fire: duplicated positions with different
    strengths, targets sitting on sources, on filament end points and on filament axes, a zero-length
    filament, a target at the origin (= the padding records), sub-underflow separations, one NaN and
    one inf coordinate."""
    base = "P3D_M2M_vel" if op == "P3D_M2M_vort" else op
    src, tgt = make_case(base, rng, n, m, self_targets=True)
    d = 2 if op.startswith("P2D") else 3
    src[5, :d] = src[4, :d]                      # two sources at one position, different strengths
    src[300, :d] = src[4, :d]                    # ... and a third one in the next chain
    if op.startswith("F3D"):
        src[7, 3:6] = src[7, 0:3]                # zero-length filament
        tgt[0, :3] = src[9, 0:3]                 # target on a start point
        tgt[1, :3] = src[9, 3:6]                 # target on an end point
        tgt[2, :3] = 0.5 * (src[11, 0:3] + src[11, 3:6])          # on the segment
        tgt[3, :3] = src[12, 0:3] + 3.0 * (src[12, 3:6] - src[12, 0:3])   # on the axis, outside
        tgt[4, :3] = (0.37, 0.0, 0.0)            # on the padding filament's axis
    else:
        tgt[0, :d] = src[4, :d]
        tgt[1, :d] = src[400, :d]
        tgt[2, :d] = src[n - 1, :d]
        tgt[6, :d] = src[20, :d] + 1e-30         # separation whose square underflows
    tgt[5, :d] = 0.0                             # the origin: coincides with the particle padding
    if nonfinite:                                # (the reference turns every target into NaN for these)
        src[30, 0] = np.nan
        src[600, 1] = np.inf
    return src, tgt


LINE_DIRECTIONS = ((1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 2, 0), (0.6, 0.8, 0), (0.3, 0.5, 0.81), (0.1, 0.9, 0.4))


def vortex_line(direction, n=40, as_particles=False):
    """A straight vortex line cut into n filaments (nodes p0 + k h d evaluated in FP32) and the points
    a filament code evaluates on it: every node and every segment midpoint.  All of them lie on the
    axes of all n filaments, where |r1 x r2| is zero or rounding noise and the reference's result is
    decided by how its cross product rounds (cvortex_b200/csrc/pair_math.cuh, cross_rounded)."""
    d = np.asarray(direction, np.float32)
    p0, h = np.float32([0.3, 0.7, 0.2]), np.float32(0.125)
    nodes = np.stack([p0 + np.float32(k) * h * d for k in range(n + 1)]).astype(np.float32)
    fil = np.zeros((n, 7), np.float32)
    fil[:, 0:3], fil[:, 3:6], fil[:, 6] = nodes[:-1], nodes[1:], 1.0
    pts = np.concatenate([nodes, (0.5 * (nodes[:-1] + nodes[1:])).astype(np.float32)]).astype(np.float32)
    if as_particles:
        tgt = np.zeros((len(pts), 7), np.float32)
        tgt[:, :3], tgt[:, 3:6], tgt[:, 6] = pts, np.float32([0.2, -0.4, 0.9]), 0.01
        return fil, tgt
    return fil, pts


def vortex_ring_case(op, n=3000):
    """A closed ring of n short filaments (radius 2 in the box of 10) and three sets of points: around the core,
    in the ring's plane, 100 diameters away.  Returns (filaments, {name: targets})."""
    phi = np.linspace(0.0, 2.0 * np.pi, n + 1)
    nodes = np.stack([5.0 + 2.0 * np.cos(phi), 5.0 + 2.0 * np.sin(phi), np.full_like(phi, 5.0)], axis=1).astype(np.float32)
    fil = np.zeros((n, 7), np.float32)
    fil[:, 0:3], fil[:, 3:6], fil[:, 6] = nodes[:-1], nodes[1:], 1.3
    rng = np.random.default_rng(3)
    sets = {"near": nodes[rng.integers(0, n, 300)] + rng.normal(0.0, 0.05, (300, 3)).astype(np.float32),
            "plane": np.stack([rng.uniform(0, 10, 400), rng.uniform(0, 10, 400), np.full(400, 5.0)], axis=1),
            "far": rng.uniform(-200, 200, (300, 3))}
    out = {}
    for name, pts in sets.items():
        pts = np.ascontiguousarray(pts, np.float32)
        if op.endswith("dvort"):
            pts = np.concatenate([pts, np.tile(np.float32([[0.2, -0.4, 0.9, 0.01]]), (len(pts), 1))], axis=1)
        out[name] = np.ascontiguousarray(pts, np.float32)
    return fil, out


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


# ---- redistribution onto a grid ------------------------------------------------------------
REDISTS = ("lambda0", "lambda1", "lambda2", "lambda3", "m4p")


def remesh_particles(rng, n, dim, box=1.0, signed=True):
    """Particles for a redistribution case: coords uniform in [0, box)^dim (the reference's
    benchmark recipe, bench/benchredistribution.c:49 + bench/bencharraysetup.c), vorticity
    uniform in [-1, 1) (signed, so node sums cancel) or [0, 1)."""
    lo = -1.0 if signed else 0.0
    if dim == 3:
        p = np.zeros((n, 7), np.float32)
        p[:, :3] = rng.uniform(0.0, box, (n, 3))
        p[:, 3:6] = rng.uniform(lo, 1.0, (n, 3))
        p[:, 6] = 0.01
    else:
        p = np.zeros((n, 4), np.float32)
        p[:, :2] = rng.uniform(0.0, box, (n, 2))
        p[:, 2] = rng.uniform(lo, 1.0, n)
        p[:, 3] = 0.01
    return p


def remesh_cases():
    """(dim, interpolant, grid spacing, negligible_vort, max_output or None) of the golden set."""
    cases = []
    for dim, h in ((3, 0.08), (2, 0.03)):
        for name in REDISTS:
            for negl in (0.0, 0.1):
                cases.append((dim, name, h, negl, None))
        for name, cap in (("lambda1", 200), ("m4p", 150)):
            cases.append((dim, name, h, 0.01, cap))
    return cases


def match_particles(a, b):
    """Pair the particles of two redistribution results by (bit-exact) position.
    Returns (rows of a, rows of b) of the common nodes, and the unmatched rows of each."""
    dim = 3 if a.shape[1] == 7 else 2
    ka = {a[i, :dim].tobytes(): i for i in range(len(a))}
    kb = {b[i, :dim].tobytes(): i for i in range(len(b))}
    common = [k for k in ka if k in kb]
    ia = np.array([ka[k] for k in common], dtype=int)
    ib = np.array([kb[k] for k in common], dtype=int)
    only_a = a[[ka[k] for k in ka if k not in kb]] if len(ka) > len(common) else a[:0]
    only_b = b[[kb[k] for k in kb if k not in ka]] if len(kb) > len(common) else b[:0]
    return a[ia], b[ib], only_a, only_b


def assert_same_remesh(got, want, tol=2e-6, what=""):
    """Two redistribution results agree: same nodes in the same order with strengths within
    `tol` of the largest strength.  Node sums are formed in a different order on every
    implementation (the reference's own depends on its thread count), so a node sitting on the
    pruning threshold may fall on either side: up to two such nodes may differ, provided they
    are among the weakest kept."""
    dim = 3 if want.shape[1] == 7 else 2
    w = slice(dim, dim + (3 if dim == 3 else 1))
    scale = float(np.abs(want[:, w]).max()) if len(want) else 1.0
    if len(got) == len(want) and np.array_equal(got[:, :dim], want[:, :dim]):
        assert np.abs(got[:, w] - want[:, w]).max(initial=0.0) <= tol * scale, what
        assert np.array_equal(got[:, -1], want[:, -1]), what
        return
    ca, cb, oa, ob = match_particles(got, want)
    assert len(oa) + len(ob) <= 2, (what, len(got), len(want), len(oa), len(ob))
    strength = lambda r: np.linalg.norm(r[:, w].astype(np.float64), axis=1)
    floor = np.sort(strength(want))[min(len(want) - 1, 3)] * 1.001
    assert all(strength(x).max(initial=0.0) <= floor for x in (oa, ob)), (what, "a strong node differs")
    assert np.abs(ca[:, w] - cb[:, w]).max(initial=0.0) <= 50 * tol * scale, what
