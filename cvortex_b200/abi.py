"""ctypes view of the ``cvtx_*`` C ABI (reference include/cvortex/libcvtx.h:53-379).

:class:`CvtxLibrary` binds *any* shared library that exports that ABI -- the
product ``libcvortex.so`` built from ``cvortex_b200/csrc`` and, in the tests,
the reference's own CPU build -- so a parity test drives both sides through
byte-identical calls: arrays of pointers to 28-/16-byte POD structs in, packed
``bsv_V3f`` / ``bsv_V2f`` / ``float`` arrays out, ``cvtx_VortFunc`` passed by
pointer after being returned *by value* from the library's own constructor.

The numpy front end mirrors the reference's function names one to one
(``P3D_M2M_vel`` = ``cvtx_P3D_M2M_vel`` ...).  Particles are rows of a float32
matrix whose row *is* the C struct:

* ``cvtx_P3D``  = ``[x, y, z, wx, wy, wz, vol]``          (libcvtx.h:53-57)
* ``cvtx_F3D``  = ``[ax, ay, az, bx, by, bz, strength]``  (libcvtx.h:60-63)
* ``cvtx_P2D``  = ``[x, y, vorticity, area]``             (libcvtx.h:66-70)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

REGULARISATIONS = ("singular", "winckelmans", "planetary", "gaussian")


class VortFunc(C.Structure):
    """``cvtx_VortFunc`` (libcvtx.h:86-94): six function pointers + a 32-char key."""
    _fields_ = [
        ("g_3D", C.CFUNCTYPE(C.c_float, C.c_float)),
        ("g_2D", C.CFUNCTYPE(C.c_float, C.c_float)),
        ("zeta_3D", C.CFUNCTYPE(C.c_float, C.c_float)),
        ("combined_3D", C.CFUNCTYPE(None, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float))),
        ("eta_3D", C.CFUNCTYPE(C.c_float, C.c_float)),
        ("eta_2D", C.CFUNCTYPE(C.c_float, C.c_float)),
        ("cl_kernel_name_ext", C.c_char * 32),
    ]


assert C.sizeof(VortFunc) == 80 and VortFunc.cl_kernel_name_ext.offset == 48

REDISTRIBUTIONS = ("lambda0", "lambda1", "lambda2", "lambda3", "m4p")


class RedistFunc(C.Structure):
    """``cvtx_RedistFunc`` (libcvtx.h:96-99): the 1-D interpolant and its support radius in cells."""
    _fields_ = [("func", C.CFUNCTYPE(C.c_float, C.c_float)), ("radius", C.c_float)]


class V3f(C.Structure):
    _fields_ = [("x", C.c_float * 3)]


class V2f(C.Structure):
    _fields_ = [("x", C.c_float * 2)]


_vp = C.c_void_p
_VFp = C.POINTER(VortFunc)


def _rows(a, cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected an (n, {cols}) float32 array, got {a.shape}")
    return a


def pointer_array(rows: np.ndarray) -> np.ndarray:
    """The ``const cvtx_P3D **`` the ABI wants: one pointer per row of ``rows``."""
    stride = rows.strides[0]
    return rows.ctypes.data + stride * np.arange(rows.shape[0], dtype=np.uint64)


class PointerRows:
    """Particles as a C caller holds them: the struct array plus the array of pointers to its
    elements, built once (the reference's benchmark does the same in its setup,
    bench/bencharraysetup.c:43-58) and passed to any number of calls."""

    def __init__(self, rows, cols: int):
        self.rows = _rows(rows, cols)
        self.ptrs = pointer_array(self.rows)

    @property
    def shape(self):
        return self.rows.shape


class CvtxLibrary:
    """A loaded library exporting the reference's ``cvtx_*`` ABI."""

    def __init__(self, path: str, mode: int = C.RTLD_LOCAL):
        self.path = path
        self.lib = lib = C.CDLL(path, mode=mode)
        i, f = C.c_int, C.c_float
        sig = {
            # accelerator control, libcvtx.h:102-110
            "cvtx_initialise": (None, []), "cvtx_finalise": (None, []),
            "cvtx_information": (C.c_char_p, []),
            "cvtx_num_accelerators": (i, []), "cvtx_num_enabled_accelerators": (i, []),
            "cvtx_accelerator_name": (C.c_char_p, [i]), "cvtx_accelerator_enabled": (i, [i]),
            "cvtx_accelerator_enable": (None, [i]), "cvtx_accelerator_disable": (None, [i]),
            # single pair, libcvtx.h:126-149, 267-273, 308-343
            "cvtx_P3D_S2S_vel": (V3f, [_vp, V3f, _VFp, f]),
            "cvtx_P3D_S2S_dvort": (V3f, [_vp, _vp, _VFp, f]),
            "cvtx_P3D_S2S_visc_dvort": (V3f, [_vp, _vp, _VFp, f, f]),
            "cvtx_P2D_S2S_vel": (V2f, [_vp, V2f, _VFp, f]),
            "cvtx_P2D_S2S_visc_dvort": (f, [_vp, _vp, _VFp, f, f]),
            "cvtx_F3D_S2S_vel": (V3f, [_vp, V3f]),
            "cvtx_F3D_S2S_dvort": (V3f, [_vp, _vp]),
            # many sources on one target, libcvtx.h:151-211, 275-283, 318-360
            "cvtx_P3D_M2S_vel": (V3f, [_vp, i, V3f, _VFp, f]),
            "cvtx_P3D_M2S_dvort": (V3f, [_vp, i, _vp, _VFp, f]),
            "cvtx_P3D_M2S_visc_dvort": (V3f, [_vp, i, _vp, _VFp, f, f]),
            "cvtx_P3D_M2S_vort": (V3f, [_vp, i, V3f, _VFp, f]),
            "cvtx_P2D_M2S_vel": (V2f, [_vp, i, V2f, _VFp, f]),
            "cvtx_P2D_M2S_visc_dvort": (f, [_vp, i, _vp, _VFp, f, f]),
            "cvtx_F3D_M2S_vel": (V3f, [_vp, i, V3f]),
            "cvtx_F3D_M2S_dvort": (V3f, [_vp, i, _vp]),
            # the hot path, libcvtx.h:213-248, 285-297, 329-336, 362-370
            "cvtx_P3D_M2M_vel": (None, [_vp, i, _vp, i, _vp, _VFp, f]),
            "cvtx_P3D_M2M_dvort": (None, [_vp, i, _vp, i, _vp, _VFp, f]),
            "cvtx_P3D_M2M_visc_dvort": (None, [_vp, i, _vp, i, _vp, _VFp, f, f]),
            "cvtx_P3D_M2M_vort": (None, [_vp, i, _vp, i, _vp, _VFp, f]),
            "cvtx_P2D_M2M_vel": (None, [_vp, i, _vp, i, _vp, _VFp, f]),
            "cvtx_P2D_M2M_visc_dvort": (None, [_vp, i, _vp, i, _vp, _VFp, f, f]),
            "cvtx_F3D_M2M_vel": (None, [_vp, i, _vp, i, _vp]),
            "cvtx_F3D_M2M_dvort": (None, [_vp, i, _vp, i, _vp]),
            "cvtx_F3D_inf_mtrx": (None, [_vp, i, _vp, _vp, i, _vp]),          # libcvtx.h:299-305
            # remeshing / relaxation, libcvtx.h:250-265, 372-379
            "cvtx_P3D_redistribute_on_grid": (i, [_vp, i, _vp, i, C.POINTER(RedistFunc), f, f]),
            "cvtx_P2D_redistribute_on_grid": (i, [_vp, i, _vp, i, C.POINTER(RedistFunc), f, f]),
            "cvtx_P3D_pedrizzetti_relaxation": (None, [_vp, i, f, _VFp, f]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        self._vf = {}
        for reg in REGULARISATIONS:
            ctor = getattr(lib, f"cvtx_VortFunc_{reg}")
            ctor.restype, ctor.argtypes = VortFunc, []
            self._vf[reg] = ctor()
        self._rf = {}
        for name in REDISTRIBUTIONS:
            ctor = getattr(lib, f"cvtx_RedistFunc_{name}")
            ctor.restype, ctor.argtypes = RedistFunc, []
            self._rf[name] = ctor()

    # ---- lifecycle / accelerators (reference src/accelerators.cpp:39-118) ----
    def initialise(self): self.lib.cvtx_initialise()
    def finalise(self): self.lib.cvtx_finalise()
    def information(self) -> str: return (self.lib.cvtx_information() or b"").decode()
    def num_accelerators(self) -> int: return self.lib.cvtx_num_accelerators()
    def num_enabled_accelerators(self) -> int: return self.lib.cvtx_num_enabled_accelerators()
    def accelerator_enabled(self, k: int) -> int: return self.lib.cvtx_accelerator_enabled(k)
    def accelerator_enable(self, k: int): self.lib.cvtx_accelerator_enable(k)
    def accelerator_disable(self, k: int): self.lib.cvtx_accelerator_disable(k)

    def accelerator_name(self, k: int):
        s = self.lib.cvtx_accelerator_name(k)
        return None if s is None else s.decode()

    def vortfunc(self, reg) -> VortFunc:
        """``cvtx_VortFunc_<reg>()`` of *this* library (or pass a VortFunc through)."""
        return reg if isinstance(reg, VortFunc) else self._vf[reg]

    def redistfunc(self, name) -> RedistFunc:
        """``cvtx_RedistFunc_<name>()`` of *this* library (or pass a RedistFunc through)."""
        return name if isinstance(name, RedistFunc) else self._rf[name]

    # ---- single pair ----
    def P3D_S2S_vel(self, p, x, reg, sigma):
        p = _rows(np.atleast_2d(p), 7)
        r = self.lib.cvtx_P3D_S2S_vel(p.ctypes.data, V3f((C.c_float * 3)(*map(float, x))),
                                      C.byref(self.vortfunc(reg)), sigma)
        return np.array(r.x[:], dtype=np.float32)

    def P3D_S2S_dvort(self, p, q, reg, sigma):
        p, q = _rows(np.atleast_2d(p), 7), _rows(np.atleast_2d(q), 7)
        r = self.lib.cvtx_P3D_S2S_dvort(p.ctypes.data, q.ctypes.data, C.byref(self.vortfunc(reg)), sigma)
        return np.array(r.x[:], dtype=np.float32)

    def P3D_S2S_visc_dvort(self, p, q, reg, sigma, nu):
        p, q = _rows(np.atleast_2d(p), 7), _rows(np.atleast_2d(q), 7)
        r = self.lib.cvtx_P3D_S2S_visc_dvort(p.ctypes.data, q.ctypes.data,
                                             C.byref(self.vortfunc(reg)), sigma, nu)
        return np.array(r.x[:], dtype=np.float32)

    def P2D_S2S_vel(self, p, x, reg, sigma):
        p = _rows(np.atleast_2d(p), 4)
        r = self.lib.cvtx_P2D_S2S_vel(p.ctypes.data, V2f((C.c_float * 2)(*map(float, x))),
                                      C.byref(self.vortfunc(reg)), sigma)
        return np.array(r.x[:], dtype=np.float32)

    def P2D_S2S_visc_dvort(self, p, q, reg, sigma, nu):
        p, q = _rows(np.atleast_2d(p), 4), _rows(np.atleast_2d(q), 4)
        return float(self.lib.cvtx_P2D_S2S_visc_dvort(p.ctypes.data, q.ctypes.data,
                                                      C.byref(self.vortfunc(reg)), sigma, nu))

    def F3D_S2S_vel(self, fil, x):
        fil = _rows(np.atleast_2d(fil), 7)
        r = self.lib.cvtx_F3D_S2S_vel(fil.ctypes.data, V3f((C.c_float * 3)(*map(float, x))))
        return np.array(r.x[:], dtype=np.float32)

    def F3D_S2S_dvort(self, fil, q):
        fil, q = _rows(np.atleast_2d(fil), 7), _rows(np.atleast_2d(q), 7)
        r = self.lib.cvtx_F3D_S2S_dvort(fil.ctypes.data, q.ctypes.data)
        return np.array(r.x[:], dtype=np.float32)

    # ---- many sources on one target (M2S) ----
    def M2S(self, op: str, src, tgt_row, reg="singular", sigma=1.0, nu=0.0):
        """cvtx_<op with M2M replaced by M2S>: `src` rows on ONE target row (a point, or a particle row for
        the dvort ops); returns the output row as a float32 array."""
        name = "cvtx_" + op.replace("M2M", "M2S")
        two_d, fil = op.startswith("P2D"), op.startswith("F3D")
        src = src if isinstance(src, PointerRows) else PointerRows(src, 4 if two_d else 7)
        particle_target = op.endswith("dvort")
        if particle_target:
            row = _rows(np.atleast_2d(tgt_row), 4 if two_d else 7)
            targ = row.ctypes.data
        elif two_d:
            targ = V2f((C.c_float * 2)(*map(float, tgt_row)))
        else:
            targ = V3f((C.c_float * 3)(*map(float, tgt_row)))
        tail = () if fil else ((C.byref(self.vortfunc(reg)), sigma, nu) if op.endswith("visc_dvort") else (C.byref(self.vortfunc(reg)), sigma))
        r = getattr(self.lib, name)(src.ptrs.ctypes.data, src.shape[0], targ, *tail)
        return np.array([r], dtype=np.float32) if isinstance(r, float) else np.array(r.x[:], dtype=np.float32)

    # ---- the hot path: all-pairs M2M ----
    def _m2m(self, name, src, scols, tgt, tcols, ocols, tail, tgt_is_particles, out=None):
        src = src if isinstance(src, PointerRows) else PointerRows(src, scols)
        if out is None:
            out = np.full((tgt.shape[0], ocols), np.nan, dtype=np.float32)
        if tgt_is_particles:
            tgt = tgt if isinstance(tgt, PointerRows) else PointerRows(tgt, tcols)
            targ = tgt.ptrs.ctypes.data
        else:
            tgt = _rows(tgt, tcols)
            targ = tgt.ctypes.data
        getattr(self.lib, name)(src.ptrs.ctypes.data, src.shape[0], targ, tgt.shape[0], out.ctypes.data, *tail)
        return out[:, 0] if ocols == 1 and out.ndim == 2 else out

    def P3D_M2M_vel(self, particles, mes, reg, sigma, out=None):
        """cvtx_P3D_M2M_vel (libcvtx.h:213-220): (n,7) particles on (m,3) points -> (m,3)."""
        return self._m2m("cvtx_P3D_M2M_vel", particles, 7, mes, 3, 3,
                         (C.byref(self.vortfunc(reg)), sigma), False, out)

    def P3D_M2M_dvort(self, particles, induced, reg, sigma, out=None):
        """cvtx_P3D_M2M_dvort (libcvtx.h:222-229): (n,7) on (m,7) particles -> (m,3)."""
        return self._m2m("cvtx_P3D_M2M_dvort", particles, 7, induced, 7, 3,
                         (C.byref(self.vortfunc(reg)), sigma), True, out)

    def P3D_M2M_visc_dvort(self, particles, induced, reg, sigma, nu, out=None):
        """cvtx_P3D_M2M_visc_dvort (libcvtx.h:231-239)."""
        return self._m2m("cvtx_P3D_M2M_visc_dvort", particles, 7, induced, 7, 3,
                         (C.byref(self.vortfunc(reg)), sigma, nu), True, out)

    def P3D_M2M_vort(self, particles, mes, reg, sigma, out=None):
        """cvtx_P3D_M2M_vort (libcvtx.h:241-248)."""
        return self._m2m("cvtx_P3D_M2M_vort", particles, 7, mes, 3, 3,
                         (C.byref(self.vortfunc(reg)), sigma), False, out)

    def P2D_M2M_vel(self, particles, mes, reg, sigma, out=None):
        """cvtx_P2D_M2M_vel (libcvtx.h:329-336): (n,4) on (m,2) -> (m,2)."""
        return self._m2m("cvtx_P2D_M2M_vel", particles, 4, mes, 2, 2,
                         (C.byref(self.vortfunc(reg)), sigma), False, out)

    def P2D_M2M_visc_dvort(self, particles, induced, reg, sigma, nu, out=None):
        """cvtx_P2D_M2M_visc_dvort (libcvtx.h:362-370): (n,4) on (m,4) -> (m,)."""
        return self._m2m("cvtx_P2D_M2M_visc_dvort", particles, 4, induced, 4, 1,
                         (C.byref(self.vortfunc(reg)), sigma, nu), True, out)

    def F3D_M2M_vel(self, filaments, mes, out=None):
        """cvtx_F3D_M2M_vel (libcvtx.h:285-290): (n,7) filaments on (m,3) -> (m,3)."""
        return self._m2m("cvtx_F3D_M2M_vel", filaments, 7, mes, 3, 3, (), False, out)

    def F3D_M2M_dvort(self, filaments, induced, out=None):
        """cvtx_F3D_M2M_dvort (libcvtx.h:292-297): (n,7) filaments on (m,7) particles -> (m,3)."""
        return self._m2m("cvtx_F3D_M2M_dvort", filaments, 7, induced, 7, 3, (), True, out)

    def F3D_inf_mtrx(self, filaments, mes, dirs, out=None):
        """cvtx_F3D_inf_mtrx (libcvtx.h:299-305): (m, n) matrix, [i, j] = u_j(mes_i) . dir_i."""
        fil = filaments if isinstance(filaments, PointerRows) else PointerRows(filaments, 7)
        mes, dirs = _rows(mes, 3), _rows(dirs, 3)
        if out is None:
            out = np.full((mes.shape[0], fil.shape[0]), np.nan, dtype=np.float32)
        self.lib.cvtx_F3D_inf_mtrx(fil.ptrs.ctypes.data, fil.shape[0], mes.ctypes.data, dirs.ctypes.data,
                                   mes.shape[0], out.ctypes.data)
        return out


    # ---- remeshing and relaxation (the steps either side of the hot path in a time step) ----
    def _redistribute(self, name, particles, cols, redist, grid_density, negligible_vort, max_output, count_only, out=None):
        src = particles if isinstance(particles, PointerRows) else PointerRows(particles, cols)
        rf = self.redistfunc(redist)
        fn = getattr(self.lib, name)
        if count_only:
            return fn(src.ptrs.ctypes.data, src.shape[0], None, 0, C.byref(rf), grid_density, negligible_vort)
        if max_output is None:   # the reference's own idiom: ask for the count, then for the particles
            max_output = fn(src.ptrs.ctypes.data, src.shape[0], None, 0, C.byref(rf), grid_density, negligible_vort)
        if out is None:
            out = np.full((max(max_output, 1), cols), np.nan, dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape[0] >= max_output and out.shape[1] == cols
        n = fn(src.ptrs.ctypes.data, src.shape[0], out.ctypes.data, max_output, C.byref(rf), grid_density, negligible_vort)
        return out[:n]

    def P3D_redistribute_on_grid(self, particles, redist, grid_density, negligible_vort=0.0, max_output=None,
                                 count_only=False, out=None):
        """cvtx_P3D_redistribute_on_grid (libcvtx.h:250-257): (n,7) particles -> (k,7) particles on grid nodes.
        `out`: a caller-owned (>= max_output, 7) float32 array to write into, as a C caller would hold."""
        return self._redistribute("cvtx_P3D_redistribute_on_grid", particles, 7, redist, grid_density,
                                  negligible_vort, max_output, count_only, out)

    def P2D_redistribute_on_grid(self, particles, redist, grid_density, negligible_vort=0.0, max_output=None,
                                 count_only=False, out=None):
        """cvtx_P2D_redistribute_on_grid (libcvtx.h:372-379): (n,4) particles -> (k,4) particles on grid nodes."""
        return self._redistribute("cvtx_P2D_redistribute_on_grid", particles, 4, redist, grid_density,
                                  negligible_vort, max_output, count_only, out)

    def P3D_pedrizzetti_relaxation(self, particles, fdt, reg, sigma):
        """cvtx_P3D_pedrizzetti_relaxation (libcvtx.h:259-265): returns the (n,7) particles with relaxed vorticity."""
        src = PointerRows(np.array(particles, dtype=np.float32, copy=True), 7)
        self.lib.cvtx_P3D_pedrizzetti_relaxation(src.ptrs.ctypes.data, src.shape[0], fdt,
                                                 C.byref(self.vortfunc(reg)), sigma)
        return src.rows
