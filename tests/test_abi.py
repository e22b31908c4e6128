"""The drop-in boundary: libcvortex.so loads, exports exactly what include/*.h declare, keeps
the reference's struct layouts, and behaves like the reference where no GPU is involved."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from util import call_abi, make_case, op_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", text)))


def _exported():
    from cvortex_b200 import _native
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    return sorted(line.split()[-1] for line in out.splitlines() if " T " in line)


def test_exports_match_headers(product):
    public = [s for s in _declared("cvortex/libcvtx.h", "cvtx_") if not s.startswith("cvtx_b200")]
    thin = _declared("cvtx_b200.h", "cvtx_b200_")
    assert len(public) == 52, len(public)
    exported = _exported()
    assert sorted(public + thin) == exported, set(exported) ^ set(public + thin)
    for name in public + thin:                       # and each one resolves through the loader
        assert getattr(product.lib, name) is not None


def test_exports_match_reference_library(ref):
    if ref is None:
        pytest.skip("oracle/_ref not built on this box")
    out = subprocess.run(["nm", "-D", "--defined-only", ref.path], capture_output=True, text=True, check=True).stdout
    theirs = sorted(line.split()[-1] for line in out.splitlines() if " T cvtx_" in line)
    ours = [s for s in _exported() if not s.startswith("cvtx_b200")]
    assert theirs == ours


def test_struct_layouts():
    from cvortex_b200.abi import V2f, V3f, VortFunc
    assert C.sizeof(V3f) == 12 and C.sizeof(V2f) == 8
    assert C.sizeof(VortFunc) == 80 and VortFunc.cl_kernel_name_ext.offset == 48
    src = r'''
    #include <stddef.h>
    #include <cvortex/libcvtx.h>
    #ifdef __cplusplus
    #define _Static_assert static_assert
    #endif
    _Static_assert(sizeof(cvtx_P3D) == 28 && sizeof(cvtx_F3D) == 28 && sizeof(cvtx_P2D) == 16, "particles");
    _Static_assert(sizeof(bsv_V3f) == 12 && sizeof(bsv_V2f) == 8 && sizeof(bsv_V3d) == 24, "bsv");
    _Static_assert(sizeof(cvtx_VortFunc) == 80 && offsetof(cvtx_VortFunc, cl_kernel_name_ext) == 48, "vortfunc");
    int main(void) { bsv_V3f a = {{1, 2, 3}}, b = {{4, 5, 6}}; return bsv_V3f_dot(a, b) == 32.f ? 0 : 1; }
    '''
    for cc, std in (("/usr/bin/gcc", "-std=c99"), ("/usr/bin/g++", "-std=c++17")):   # the header is C and C++
        res = subprocess.run([cc, std, "-x", "c" if cc.endswith("gcc") else "c++", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), "-"],
                             input=src, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_lifecycle_and_accelerator_api_without_assuming_a_gpu(product):
    n = product.num_accelerators()
    assert n >= 0
    assert "cvortex version: 0.3.8" in product.information() and "CUDA" in product.information()
    assert product.accelerator_name(n) is None and product.accelerator_name(-1) is None
    assert product.accelerator_enabled(n) == 0
    if n:
        assert product.accelerator_name(0)
        assert product.accelerator_enabled(0) == 1          # default: device 0, as the reference
    product.finalise()                                       # finalise -> initialise must work
    product.initialise()                                     # (reference bench/benchinitilisation.h:9-14)
    product.initialise()                                     # and is idempotent
    assert product.num_accelerators() == n


@pytest.mark.parametrize("op,reg", op_cases() + [("P3D_M2M_vort", "gaussian")])
def test_host_loops_equal_reference_when_accelerators_are_disabled(product, oracle, op, reg):
    """disable(all) is the reference's CPU switch; the host loops behind it reproduce the
    reference's OpenMP path bit for bit (oracle == reference, see test_oracle.py)."""
    from cvortex_b200.device import DeviceBackend
    dev = DeviceBackend(product.lib)
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    try:
        rng = np.random.default_rng(3)
        src, tgt = make_case(op if op != "P3D_M2M_vort" else "P3D_M2M_vel", rng, 150, 60, self_targets=True)
        got = call_abi(product, op, src, tgt, reg, 0.3, 0.1)
        assert dev.last_dispatch() == 0
        assert np.array_equal(got, oracle.m2m(op, src, tgt, reg, 0.3, 0.1))
    finally:
        for k in enabled:
            product.accelerator_enable(k)


def test_scalar_entry_points_match_oracle(product, oracle):
    rng = np.random.default_rng(5)
    p, q = rng.uniform(0, 2, 7).astype(np.float32), rng.uniform(0, 2, 7).astype(np.float32)
    p2, q2 = rng.uniform(0, 2, 4).astype(np.float32), rng.uniform(0, 2, 4).astype(np.float32)
    for reg in ("singular", "winckelmans", "planetary", "gaussian"):
        assert np.array_equal(product.P3D_S2S_vel(p, q[:3], reg, 0.4), oracle.s2s("P3D_S2S_vel", p, q[:3], reg, 0.4))
        assert np.array_equal(product.P3D_S2S_dvort(p, q, reg, 0.4), oracle.s2s("P3D_S2S_dvort", p, q, reg, 0.4))
        assert np.array_equal(product.P2D_S2S_vel(p2, q2[:2], reg, 0.4), oracle.s2s("P2D_S2S_vel", p2, q2[:2], reg, 0.4))
    for reg in ("winckelmans", "gaussian"):
        assert np.array_equal(product.P3D_S2S_visc_dvort(p, q, reg, 0.4, 0.2), oracle.s2s("P3D_S2S_visc_dvort", p, q, reg, 0.4, 0.2))
        assert product.P2D_S2S_visc_dvort(p2, q2, reg, 0.4, 0.2) == oracle.s2s("P2D_S2S_visc_dvort", p2, q2, reg, 0.4, 0.2)[0]
    assert np.array_equal(product.F3D_S2S_vel(p, q[:3]), oracle.s2s("F3D_S2S_vel", p, q[:3]))
    assert np.array_equal(product.F3D_S2S_dvort(p, q), oracle.s2s("F3D_S2S_dvort", p, q))


def test_thin_abi_rejects_bad_arguments_without_a_gpu(product):
    from cvortex_b200.device import BackendError, DeviceBackend
    dev = DeviceBackend(product.lib)
    info = dev.op_info("P3D_M2M_vel", "winckelmans")
    assert info == {"src_cols": 7, "tgt_cols": 3, "out_cols": 3, "lane_ops": 21, "sfu_ops": 1}
    assert dev.op_info("P2D_M2M_visc_dvort", "gaussian")["out_cols"] == 1
    with pytest.raises(BackendError):
        dev.op_info("P3D_M2M_visc_dvort", "planetary")       # no eta for planetary (reference src/VortFunc.cpp:232)
    with pytest.raises(BackendError):                          # a device that does not exist is an error, never a CPU run
        dev.m2m_host("P3D_M2M_vel", "winckelmans", 1000, np.zeros((2, 7), np.float32), np.zeros((2, 3), np.float32))


def test_python_mirror_refuses_to_run_without_a_gpu(product):
    from cvortex_b200 import api
    from cvortex_b200.device import BackendError
    if product.num_accelerators() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(BackendError):
        api.initialise(require_gpu=True)
    with pytest.raises(BackendError):
        api.P3D_M2M_vel(np.zeros((4, 7), np.float32), np.zeros((4, 3), np.float32), "winckelmans", 0.1)


def test_inf_mtrx_host_path_equals_reference(product, oracle):
    """cvtx_F3D_inf_mtrx with every accelerator disabled: the reference's loops, bit for bit."""
    from util import filaments, points
    enabled = [k for k in range(product.num_accelerators()) if product.accelerator_enabled(k)]
    for k in enabled:
        product.accelerator_disable(k)
    try:
        rng = np.random.default_rng(2)
        F, X = filaments(rng, 90, seg=0.5), points(rng, 37, 3)
        D = rng.uniform(-1, 1, (37, 3)).astype(np.float32)
        assert np.array_equal(product.F3D_inf_mtrx(F, X, D), oracle.inf_mtrx(F, X, D))
    finally:
        for k in enabled:
            product.accelerator_enable(k)
