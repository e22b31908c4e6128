"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle port (oracle/cvtx_oracle.c) -- the checker, never the product."""
    from oracle import binding
    binding.build(ref=True)
    return binding.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own OpenMP CPU path compiled from /root/reference (oracle/_ref), or None."""
    from oracle import binding
    from cvortex_b200.abi import CvtxLibrary
    if not binding.have_ref():
        if os.path.isdir("/root/reference/src"):
            binding.build(ref=True)
        else:
            return None
    lib = CvtxLibrary(binding.REF_SO)
    lib.initialise()
    return lib


@pytest.fixture(scope="session")
def product():
    """libcvortex.so through the cvtx_* ABI; builds it if the .so is not there yet."""
    from cvortex_b200 import _native
    from cvortex_b200.abi import CvtxLibrary
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    lib = CvtxLibrary(_native.LIB_PATH)
    lib.initialise()
    yield lib
    lib.finalise()


@pytest.fixture(scope="session")
def gpu(product):
    """The product library with a live accelerator: GPU tests fail (not skip) without one."""
    from cvortex_b200.device import DeviceBackend
    dev = DeviceBackend(product.lib)
    assert product.num_accelerators() > 0, "no CUDA accelerator: " + dev.last_error()
    for k in range(product.num_accelerators()):
        (product.accelerator_enable if k == 0 else product.accelerator_disable)(k)
    return product, dev


@pytest.fixture(scope="module")
def torch_cuda():
    """torch with a live CUDA device 0 current (device memory and streams for the thin-ABI tests)."""
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="session")
def hostcheck():
    """tests/hostcheck: the device pair arithmetic compiled for the host (test-only)."""
    import ctypes as C
    import numpy as np
    src = os.path.join(ROOT, "tests", "hostcheck", "pair_math_host.cpp")
    out = os.path.join(ROOT, "tests", "_build", "libpairmath_host.so")
    deps = [src] + [os.path.join(ROOT, "cvortex_b200", "csrc", f) for f in ("pair_math.cuh", "op_table.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-ffp-contract=off", "-shared",
                        "-o", out, src], check=True)
    lib = C.CDLL(out)
    fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    lib.hostcheck_m2m.argtypes = [C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_float, C.c_float]
    lib.hostcheck_m2m.restype = C.c_int
    lib.hostcheck_meta.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.hostcheck_meta.restype = C.c_int
    lib.hostcheck_guarded_only.argtypes, lib.hostcheck_guarded_only.restype = [C.c_int], None
    lib.hostcheck_f3d_mode.argtypes, lib.hostcheck_f3d_mode.restype = [C.c_int], None
    lib.hostcheck_last_f3d_mode.argtypes, lib.hostcheck_last_f3d_mode.restype = [], C.c_int
    lib.hostcheck_reevaluated.argtypes, lib.hostcheck_reevaluated.restype = [C.c_int], C.c_long
    lib.hostcheck_optimistic.argtypes, lib.hostcheck_optimistic.restype = [C.c_int, C.c_int], C.c_int
    return lib
