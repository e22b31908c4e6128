"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing in
``cvortex_b200/`` (the product) does.

Two libraries are bound here:

* ``oracle/_build/libcvtx_oracle.so`` -- our CPU restatement (``cvtx_oracle.c``),
  built by ``make -C oracle port``; flat numpy arrays in, numpy arrays out.
* ``oracle/_ref/libcvortex_ref.so`` -- the reference's own OpenMP CPU path,
  compiled from ``/root/reference`` by ``make -C oracle ref``.  It speaks the
  ``cvtx_*`` C ABI of the reference's ``include/cvortex/libcvtx.h``; it is driven
  through :class:`cvortex_b200.abi.CvtxLibrary`, the same binding that drives the
  product library, so both sides of a parity test go through the identical ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libcvtx_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcvortex_ref.so")

REG_IDS = {"singular": 0, "winckelmans": 1, "planetary": 2, "gaussian": 3}
REDIST_IDS = {"lambda0": 0, "lambda1": 1, "lambda2": 2, "lambda3": 3, "m4p": 4}


def build(ref: bool = True) -> None:
    """(Re)build the oracle port and, when /root/reference exists, oracle/_ref."""
    targets = ["port"]
    if ref and os.path.isdir("/root/reference/src"):
        targets.append("ref")
        # the reference's own test program linked against the product library (needs it built)
        if os.path.exists(os.path.join(os.path.dirname(HERE), "cvortex_b200", "lib", "libcvortex.so")):
            targets += ["ref_tests", "ref_bench"]
    subprocess.run(["make", "-C", HERE, "--no-print-directory"] + targets, check=True,
                   stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == cols, (a.shape, cols)
    return a


class Oracle:
    """numpy front end of libcvtx_oracle.so (see oracle/cvtx_oracle.h)."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = lib = C.CDLL(path)
        for name in ("g3d", "zeta3d", "eta3d", "g2d", "eta2d"):
            f = getattr(lib, f"cvtx_oracle_{name}")
            f.restype, f.argtypes = C.c_float, [C.c_int, C.c_float]
            f = getattr(lib, f"cvtx_oracle_{name}_f64")
            f.restype, f.argtypes = C.c_double, [C.c_int, C.c_double]
        lib.cvtx_oracle_num_threads.restype = C.c_int
        for prec, outp in (("f32", _fp), ("f64", _dp)):
            for op in ("P3D_M2M_vel", "P3D_M2M_dvort", "P3D_M2M_vort", "P2D_M2M_vel"):
                f = getattr(lib, f"cvtx_oracle_{op}_{prec}")
                f.restype, f.argtypes = None, [_fp, C.c_int, _fp, C.c_int, outp, C.c_int, C.c_float]
            for op in ("P3D_M2M_visc_dvort", "P2D_M2M_visc_dvort"):
                f = getattr(lib, f"cvtx_oracle_{op}_{prec}")
                f.restype, f.argtypes = None, [_fp, C.c_int, _fp, C.c_int, outp, C.c_int, C.c_float, C.c_float]
            for op in ("F3D_M2M_vel", "F3D_M2M_dvort"):
                f = getattr(lib, f"cvtx_oracle_{op}_{prec}")
                f.restype, f.argtypes = None, [_fp, C.c_int, _fp, C.c_int, outp]
            f = getattr(lib, f"cvtx_oracle_F3D_inf_mtrx_{prec}")
            f.restype, f.argtypes = None, [_fp, C.c_int, _fp, _fp, C.c_int, outp]
        lib.cvtx_oracle_redist.restype, lib.cvtx_oracle_redist.argtypes = C.c_float, [C.c_int, C.c_float]
        lib.cvtx_oracle_redist_radius.restype, lib.cvtx_oracle_redist_radius.argtypes = C.c_float, [C.c_int]
        for name in ("cvtx_oracle_P3D_redistribute", "cvtx_oracle_P2D_redistribute"):
            f = getattr(lib, name)
            f.restype, f.argtypes = C.c_int, [_fp, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        lib.cvtx_oracle_P3D_pedrizzetti.restype = None
        lib.cvtx_oracle_P3D_pedrizzetti.argtypes = [_fp, C.c_int, C.c_float, C.c_int, C.c_float, _fp]
        lib.cvtx_oracle_P3D_S2S_vel.argtypes = [_fp, _fp, C.c_int, C.c_float, _fp]
        lib.cvtx_oracle_P3D_S2S_dvort.argtypes = [_fp, _fp, C.c_int, C.c_float, _fp]
        lib.cvtx_oracle_P3D_S2S_visc_dvort.argtypes = [_fp, _fp, C.c_int, C.c_float, C.c_float, _fp]
        lib.cvtx_oracle_P2D_S2S_vel.argtypes = [_fp, _fp, C.c_int, C.c_float, _fp]
        lib.cvtx_oracle_P2D_S2S_visc_dvort.argtypes = [_fp, _fp, C.c_int, C.c_float, C.c_float, _fp]
        lib.cvtx_oracle_F3D_S2S_vel.argtypes = [_fp, _fp, _fp]
        lib.cvtx_oracle_F3D_S2S_dvort.argtypes = [_fp, _fp, _fp]

    # -- scalars ---------------------------------------------------------
    def scalar(self, name: str, reg: str, rho: float, f64: bool = False) -> float:
        fn = getattr(self.lib, f"cvtx_oracle_{name}" + ("_f64" if f64 else ""))
        return float(fn(REG_IDS[reg], rho))

    def num_threads(self) -> int:
        return int(self.lib.cvtx_oracle_num_threads())

    # -- single pairs ----------------------------------------------------
    def s2s(self, op: str, src, tgt, reg: str = "singular", sigma: float = 1.0, nu: float = 0.0):
        src = np.ascontiguousarray(src, dtype=np.float32).ravel()
        tgt = np.ascontiguousarray(tgt, dtype=np.float32).ravel()
        nout = {"P2D_S2S_vel": 2, "P2D_S2S_visc_dvort": 1}.get(op, 3)
        out = np.zeros(nout, dtype=np.float32)
        fn = getattr(self.lib, f"cvtx_oracle_{op}")
        if op.startswith("F3D"):
            fn(src, tgt, out)
        elif op.endswith("visc_dvort"):
            fn(src, tgt, REG_IDS[reg], sigma, nu, out)
        else:
            fn(src, tgt, REG_IDS[reg], sigma, out)
        return out

    def inf_mtrx(self, fil, pts, dirs, f64: bool = False) -> np.ndarray:
        """Dense (m, n) influence matrix of n filaments on m points along m directions."""
        fil, pts, dirs = _f32(fil, 7), _f32(pts, 3), _f32(dirs, 3)
        out = np.zeros((pts.shape[0], fil.shape[0]), dtype=np.float64 if f64 else np.float32)
        getattr(self.lib, "cvtx_oracle_F3D_inf_mtrx_" + ("f64" if f64 else "f32"))(fil, fil.shape[0], pts, dirs, pts.shape[0], out)
        return out

    # -- redistribution / relaxation -------------------------------------
    def redist(self, which: str, U: float) -> float:
        return float(self.lib.cvtx_oracle_redist(REDIST_IDS[which], U))

    def redist_radius(self, which: str) -> float:
        return float(self.lib.cvtx_oracle_redist_radius(REDIST_IDS[which]))

    def redistribute(self, particles, which: str, grid_density: float, negligible_vort: float = 0.0,
                     max_output=None, count_only: bool = False):
        """Particles (n,7) or (n,4) -> the particles the reference would create on the grid."""
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        cols = particles.shape[1]
        fn = self.lib.cvtx_oracle_P3D_redistribute if cols == 7 else self.lib.cvtx_oracle_P2D_redistribute
        n = particles.shape[0]
        dummy = np.zeros((1, cols), dtype=np.float32)
        if count_only:
            return fn(particles, n, dummy, 0, 0, REDIST_IDS[which], grid_density, negligible_vort)
        if max_output is None:
            max_output = fn(particles, n, dummy, 0, 0, REDIST_IDS[which], grid_density, negligible_vort)
        out = np.full((max(max_output, 1), cols), np.nan, dtype=np.float32)
        k = fn(particles, n, out, max_output, 1, REDIST_IDS[which], grid_density, negligible_vort)
        return out[:k]

    def pedrizzetti(self, particles, fdt: float, reg: str, sigma: float) -> np.ndarray:
        particles = _f32(particles, 7)
        out = np.empty_like(particles)
        self.lib.cvtx_oracle_P3D_pedrizzetti(particles, particles.shape[0], fdt, REG_IDS[reg], sigma, out)
        return out

    # -- M2M -------------------------------------------------------------
    _SHAPES = {  # op -> (source cols, target cols, output cols)
        "P3D_M2M_vel": (7, 3, 3), "P3D_M2M_dvort": (7, 7, 3), "P3D_M2M_visc_dvort": (7, 7, 3),
        "P3D_M2M_vort": (7, 3, 3), "P2D_M2M_vel": (4, 2, 2), "P2D_M2M_visc_dvort": (4, 4, 1),
        "F3D_M2M_vel": (7, 3, 3), "F3D_M2M_dvort": (7, 7, 3),
    }

    def m2m(self, op: str, src, tgt, reg: str = "singular", sigma: float = 1.0, nu: float = 0.0,
            f64: bool = False) -> np.ndarray:
        sc, tc, oc = self._SHAPES[op]
        src, tgt = _f32(src, sc), _f32(tgt, tc)
        out = np.zeros((tgt.shape[0], oc), dtype=np.float64 if f64 else np.float32)
        fn = getattr(self.lib, f"cvtx_oracle_{op}_" + ("f64" if f64 else "f32"))
        n, m = src.shape[0], tgt.shape[0]
        if op.startswith("F3D"):
            fn(src, n, tgt, m, out)
        elif op.endswith("visc_dvort"):
            fn(src, n, tgt, m, out, REG_IDS[reg], sigma, nu)
        else:
            fn(src, n, tgt, m, out, REG_IDS[reg], sigma)
        return out[:, 0] if oc == 1 else out
