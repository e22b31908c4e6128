"""Pin the CPU oracle (oracle/cvtx_oracle.c) to the reference before anything trusts it:
   * the 24 regularisation known-answer tests of reference test/testvortfunc.h:37-67
   * the 31 structural single-pair checks of reference test/testparticle.h:48-109
   * outputs of the reference's own OpenMP CPU path -- live when oracle/_ref was built
     from /root/reference, and always through tests/golden/reference_m2m.npz."""
import os

import numpy as np
import pytest

from util import call_abi, make_case, op_cases, rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_m2m.npz")


# ---- reference test/testvortfunc.h:37-67, value by value --------------------------------
KAT_EXACT = [("singular", "g3d", 0.1, 1.0), ("singular", "g3d", 3.0, 1.0), ("singular", "zeta3d", 0.1, 0.0),
             ("singular", "zeta3d", 10.0, 0.0), ("winckelmans", "g3d", 0.0, 0.0), ("winckelmans", "zeta3d", 0.0, 7.5),
             ("planetary", "g3d", 0.0, 0.0), ("planetary", "g3d", 1.0, 1.0), ("planetary", "g3d", 10.0, 1.0),
             ("planetary", "zeta3d", 0.99, 3.0), ("planetary", "zeta3d", 1.01, 0.0), ("gaussian", "g3d", 0.0, 0.0),
             ("gaussian", "g3d", 10.0, 1.0)]
KAT_CLOSE = [("winckelmans", "g3d", 1.0, 0.61872, 1e-5), ("winckelmans", "g3d", 10.0, 0.9998168, 1e-7),
             ("winckelmans", "zeta3d", 10.0, 7.2433e-7, 1e-11), ("gaussian", "g3d", 0.5, 0.030859595, 1e-6),
             ("gaussian", "g3d", 1.0, 0.198748043, 1e-6), ("gaussian", "g3d", 2.0, 0.738535870, 1e-6),
             ("gaussian", "g3d", 4.0, 0.998866015, 1e-6), ("gaussian", "g3d", 6.0, 0.999999925, 1e-6),
             ("gaussian", "g3d", 8.0, 0.999999999, 1e-6), ("gaussian", "zeta3d", 1.0, 0.483941449, 1e-6),
             ("gaussian", "zeta3d", 0.5, 0.70413065, 1e-6)]


def test_vortfunc_known_answers_oracle(oracle):
    assert len(KAT_EXACT) + len(KAT_CLOSE) == 24
    for reg, fn, rho, want in KAT_EXACT:
        assert oracle.scalar(fn, reg, rho) == want, (reg, fn, rho)
    for reg, fn, rho, want, tol in KAT_CLOSE:
        assert abs(oracle.scalar(fn, reg, rho) - np.float32(want)) < tol, (reg, fn, rho)
        assert abs(oracle.scalar(fn, reg, rho, f64=True) - want) < max(tol, 2e-7), (reg, fn, rho)


def test_vortfunc_known_answers_product_function_pointers(product):
    """The same 24 values through the cvtx_VortFunc tables libcvortex.so hands out."""
    for reg, fn, rho, want in KAT_EXACT:
        f = getattr(product.vortfunc(reg), {"g3d": "g_3D", "zeta3d": "zeta_3D"}[fn])
        assert f(rho) == want, (reg, fn, rho)
    for reg, fn, rho, want, tol in KAT_CLOSE:
        f = getattr(product.vortfunc(reg), {"g3d": "g_3D", "zeta3d": "zeta_3D"}[fn])
        assert abs(f(rho) - np.float32(want)) < tol, (reg, fn, rho)
    for reg in ("singular", "winckelmans", "planetary", "gaussian"):
        assert product.vortfunc(reg).cl_kernel_name_ext == reg.encode()


# ---- reference test/testparticle.h:48-109 -------------------------------------------------
def _particle_checks(vel, dvort):
    p1 = [0, 0, 0, 1, 0, 0, 1]
    v0, vx, vy, vz, vzbig = [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 100]
    n = 0
    for reg in ("winckelmans", "singular"):
        assert np.all(vel(p1, v0, reg) == 0); n += 1
        assert np.all(vel(p1, vx, reg) == 0); n += 1
        assert np.any(vel(p1, vy, reg) != 0); n += 1
        assert np.any(vel(p1, vz, reg) != 0); n += 1
    for reg in ("winckelmans", "singular"):
        a, b = vel(p1, vy, reg), vel(p1, vz, reg)
        assert a[0] == 0 and a[1] == 0 and a[2] > 0; n += 3
        assert b[0] == 0 and b[1] < 0 and b[2] == 0; n += 3
    for reg in ("winckelmans", "singular"):
        assert vel(p1, vz, reg)[1] < vel(p1, vzbig, reg)[1]; n += 1
    p2, pxz, pyz, pzz = [0, 0, 0, 0, 0, 1, 1], [0, 0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 1, 0, 1], [0, 0, 1, 0, 0, 1, 1]
    a = dvort(p2, pxz, "singular"); assert a[0] == 0 and a[1] < 0 and a[2] == 0; n += 3
    a = dvort(p2, pyz, "singular"); assert a[0] > 0 and a[1] == 0 and a[2] == 0; n += 3
    a = dvort(p2, pzz, "singular"); assert a[0] == 0 and a[1] == 0 and a[2] == 0; n += 3
    assert n == 31
    return n


def test_particle_structure_oracle(oracle):
    _particle_checks(lambda p, x, reg: oracle.s2s("P3D_S2S_vel", p, x, reg, 1.0),
                     lambda p, q, reg: oracle.s2s("P3D_S2S_dvort", p, q, reg, 1.0))


def test_particle_structure_product_scalar_api(product):
    _particle_checks(lambda p, x, reg: product.P3D_S2S_vel(p, x, reg, 1.0),
                     lambda p, q, reg: product.P3D_S2S_dvort(p, q, reg, 1.0))


# ---- against the reference's own CPU path ------------------------------------------------
def test_oracle_is_bit_exact_with_golden_reference_outputs(oracle):
    """Fixtures were produced by the unmodified reference (tests/golden/make_golden.py)."""
    z = np.load(GOLDEN)
    keys = sorted({k.rsplit("|", 1)[0] for k in z.files})
    assert len(keys) == 63
    for key in keys:
        regime, op, reg = key.split("|")
        sigma, nu = (float(v) for v in z[key + "|par"])
        got = oracle.m2m(op, z[key + "|src"], z[key + "|tgt"], reg, sigma, nu)
        want = z[key + "|out"].reshape(got.shape)
        assert np.array_equal(got, want, equal_nan=True), key
        # and the FP64 restatement tells the same story, to FP32 accuracy
        f64 = oracle.m2m(op, z[key + "|src"], z[key + "|tgt"], reg, sigma, nu, f64=True)
        if regime != "tiny" or reg != "gaussian":
            # the filament formulas cancel badly in FP32 (reference src/F3D.cpp:47-50,67-76):
            # the reference itself is ~1e-4 from FP64 on short segments
            assert rel_l2(want, f64) < (1e-3 if op.startswith("F3D") else 5e-5), key


@pytest.mark.parametrize("op,reg", op_cases())
def test_oracle_is_bit_exact_with_live_reference(oracle, ref, op, reg):
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box); golden fixtures cover it")
    rng = np.random.default_rng(17)
    for box, sigma in ((10.0, 0.3), (10.0, 0.02)):
        src, tgt = make_case(op, rng, 700, 300, box=box, self_targets=True)
        want = call_abi(ref, op, src, tgt, reg, sigma, 0.1)
        got = oracle.m2m(op, src, tgt, reg, sigma, 0.1)
        assert np.array_equal(got, want), (op, reg, box, sigma)


def test_reference_own_test_suite_passes_against_oracle_ref(ref):
    """The reference's all_tests (test/testmain.c) built against oracle/_ref passes 57/57: the
    accelerator, VortFunc and particle sections (the differential sections self-skip without
    an accelerator, reference test/testsamecpugpuresultmany.h:90)."""
    import subprocess
    import tempfile
    if ref is None or not os.path.isdir("/root/reference/test"):
        pytest.skip("needs /root/reference")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "all_tests")
        subprocess.run(["/usr/bin/gcc", "-std=gnu99", "-w", "-I/root/reference/include/cvortex", "-I" + os.path.join(root, "include"),
                        "-o", exe, "/root/reference/test/testmain.c", os.path.join(root, "oracle", "_ref", "libcvortex_ref.so"),
                        "-lm", "-Wl,-rpath," + os.path.join(root, "oracle", "_ref")], check=True)
        res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Passed 57 of 57" in res.stdout, res.stdout
