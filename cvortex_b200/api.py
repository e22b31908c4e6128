"""Host-side mirror of the reference's public interface for the all-pairs path.

Same names, argument meaning and result layout as the reference's C API
(include/cvortex/libcvtx.h:102-110, :213-248, :285-297, :329-370): every call
below goes through the *unchanged* ``cvtx_*`` C ABI of ``libcvortex.so`` --
arrays of pointers to particle structs in, packed result arrays out -- exactly
as a C or Julia (CVortex.jl) caller would.  numpy is only the container.

The one behavioural addition is ``initialise(require_gpu=True)``: the reference
quietly runs its OpenMP loops when it finds no accelerator; this mirror
refuses instead, so a box without a working CUDA path can never pass for a
GPU run.  (The C ABI itself keeps the reference's semantics: disabling every
accelerator with ``accelerator_disable`` selects the host loops explicitly.)
"""
from __future__ import annotations

from . import _native
from .abi import CvtxLibrary, REGULARISATIONS  # noqa: F401
from .device import BackendError, DeviceBackend

_lib: CvtxLibrary | None = None
_dev: DeviceBackend | None = None


def library() -> CvtxLibrary:
    """The loaded product library (raises if libcvortex.so is missing)."""
    global _lib, _dev
    if _lib is None:
        _native.load()  # clear error message when the .so is absent
        _lib = CvtxLibrary(_native.LIB_PATH)
        _dev = DeviceBackend(_lib.lib)
    return _lib


def backend() -> DeviceBackend:
    library()
    assert _dev is not None
    return _dev


def initialise(require_gpu: bool = True) -> None:
    """cvtx_initialise(); with ``require_gpu`` (default) raise unless a CUDA device is usable."""
    lib = library()
    lib.initialise()
    if require_gpu and lib.num_accelerators() <= 0:
        raise BackendError("cvtx_initialise() found no CUDA accelerator; cvortex_b200 has no CPU fallback "
                           "for the all-pairs path (pass require_gpu=False only to use the scalar host API)")


def finalise() -> None:
    library().finalise()


def information() -> str:
    return library().information()


def num_accelerators() -> int:
    return library().num_accelerators()


def num_enabled_accelerators() -> int:
    return library().num_enabled_accelerators()


def accelerator_name(k: int):
    return library().accelerator_name(k)


def accelerator_enabled(k: int) -> int:
    return library().accelerator_enabled(k)


def accelerator_enable(k: int) -> None:
    library().accelerator_enable(k)


def accelerator_disable(k: int) -> None:
    library().accelerator_disable(k)


def use_only(device: int) -> None:
    """Enable exactly one accelerator (what each rank of a one-process-per-GPU job does)."""
    lib = library()
    for k in range(lib.num_accelerators()):
        (lib.accelerator_enable if k == device else lib.accelerator_disable)(k)


def _checked(fn, *args):
    lib = library()
    if lib.num_enabled_accelerators() <= 0:
        raise BackendError("no accelerator enabled: refusing to run the all-pairs path on the host "
                           "(use library().<op> directly for the reference's explicit CPU switch)")
    res = fn(*args)
    if backend().last_dispatch() != 1:
        raise BackendError("the all-pairs call did not run on the GPU")
    return res


def P3D_M2M_vel(particles, mes, reg, sigma):
    return _checked(library().P3D_M2M_vel, particles, mes, reg, sigma)


def P3D_M2M_dvort(particles, induced, reg, sigma):
    return _checked(library().P3D_M2M_dvort, particles, induced, reg, sigma)


def P3D_M2M_visc_dvort(particles, induced, reg, sigma, nu):
    return _checked(library().P3D_M2M_visc_dvort, particles, induced, reg, sigma, nu)


def P3D_M2M_vort(particles, mes, reg, sigma):
    return _checked(library().P3D_M2M_vort, particles, mes, reg, sigma)


def P2D_M2M_vel(particles, mes, reg, sigma):
    return _checked(library().P2D_M2M_vel, particles, mes, reg, sigma)


def P2D_M2M_visc_dvort(particles, induced, reg, sigma, nu):
    return _checked(library().P2D_M2M_visc_dvort, particles, induced, reg, sigma, nu)


def F3D_M2M_vel(filaments, mes):
    return _checked(library().F3D_M2M_vel, filaments, mes)


def F3D_M2M_dvort(filaments, induced):
    return _checked(library().F3D_M2M_dvort, filaments, induced)


def F3D_inf_mtrx(filaments, mes, dirs):
    return _checked(library().F3D_inf_mtrx, filaments, mes, dirs)


# ---- the steps either side of the all-pairs sums (reference src/P3D.cpp:509-707, src/P2D.cpp:283-436) ----

def P3D_redistribute_on_grid(particles, redist, grid_density, negligible_vort=0.0, max_output=None):
    """`redist` is one of "lambda0" .. "lambda3", "m4p" (the GPU path) -- a user-defined
    cvtx_RedistFunc runs on the host: call library().P3D_redistribute_on_grid for that."""
    return _checked(library().P3D_redistribute_on_grid, particles, redist, grid_density, negligible_vort, max_output)


def P2D_redistribute_on_grid(particles, redist, grid_density, negligible_vort=0.0, max_output=None):
    return _checked(library().P2D_redistribute_on_grid, particles, redist, grid_density, negligible_vort, max_output)


def P3D_pedrizzetti_relaxation(particles, fdt, reg, sigma):
    return _checked(library().P3D_pedrizzetti_relaxation, particles, fdt, reg, sigma)
