"""The thin device-level C ABI (``include/cvtx_b200.h``) from Python.

``DeviceBackend.m2m`` takes *device pointers* (anything with ``data_ptr()`` --
torch CUDA tensors -- or raw ints) and launches on a CUDA stream; it is what a
one-process-per-GPU caller uses to keep particles resident in HBM.
``DeviceBackend.m2m_host`` takes numpy arrays and moves them itself.

Torch is plumbing here (device memory, streams): this module does not import
it, it only accepts its tensors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native

OPS = {
    "P3D_M2M_vel": 0, "P3D_M2M_dvort": 1, "P3D_M2M_visc_dvort": 2, "P3D_M2M_vort": 3,
    "P2D_M2M_vel": 4, "P2D_M2M_visc_dvort": 5, "F3D_M2M_vel": 6, "F3D_M2M_dvort": 7,
    "P3D_M2M_vel_dvort": 8,      # fused, thin ABI only (see include/cvtx_b200.h)
}
REGS = {"singular": 0, "winckelmans": 1, "planetary": 2, "gaussian": 3}
REDISTS = {"lambda0": 0, "lambda1": 1, "lambda2": 2, "lambda3": 3, "m4p": 4}


class BackendError(RuntimeError):
    pass


def _ptr(x) -> int:
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class DeviceBackend:
    """ctypes binding of the ``cvtx_b200_*`` symbols of libcvortex.so."""

    def __init__(self, lib: C.CDLL | None = None):
        self.lib = lib = lib or _native.load()
        i, f, vp, sz = C.c_int, C.c_float, C.c_void_p, C.c_size_t
        ip = C.POINTER(C.c_int)
        lib.cvtx_b200_device_count.restype = i
        lib.cvtx_b200_device_name.restype, lib.cvtx_b200_device_name.argtypes = C.c_char_p, [i]
        lib.cvtx_b200_device_sm_count.restype, lib.cvtx_b200_device_sm_count.argtypes = i, [i]
        lib.cvtx_b200_device_clock_khz.restype, lib.cvtx_b200_device_clock_khz.argtypes = i, [i]
        lib.cvtx_b200_release.restype = None
        lib.cvtx_b200_m2m.restype = i
        lib.cvtx_b200_m2m.argtypes = [i, i, i, vp, vp, i, vp, i, vp, f, f]
        lib.cvtx_b200_m2m_host.restype = i
        lib.cvtx_b200_m2m_host.argtypes = [i, i, i, vp, i, vp, i, vp, f, f, C.POINTER(sz), C.POINTER(sz)]
        lib.cvtx_b200_m2m_sharded.restype = i
        lib.cvtx_b200_m2m_sharded.argtypes = [i, i, i, ip, C.POINTER(vp), ip, C.POINTER(vp), ip, C.POINTER(vp), f, f]
        lib.cvtx_b200_exchange_backend.restype = C.c_char_p
        lib.cvtx_b200_f3d_inf_mtrx.restype = i
        lib.cvtx_b200_f3d_inf_mtrx.argtypes = [i, vp, vp, i, vp, vp, i, vp]
        lib.cvtx_b200_redistribute.restype = i
        lib.cvtx_b200_redistribute.argtypes = [i, i, i, vp, vp, i, f, f, vp, i, ip]
        lib.cvtx_b200_pedrizzetti_relaxation.restype = i
        lib.cvtx_b200_pedrizzetti_relaxation.argtypes = [i, i, vp, vp, i, f, f]
        lib.cvtx_b200_op_info.restype, lib.cvtx_b200_op_info.argtypes = i, [i, i, ip, ip, ip, ip, ip]
        lib.cvtx_b200_plan.restype, lib.cvtx_b200_plan.argtypes = i, [i, i, i, i, ip, ip, ip, ip]
        lib.cvtx_b200_kernel_launches.restype = C.c_ulonglong
        lib.cvtx_b200_measure_peak.restype, lib.cvtx_b200_measure_peak.argtypes = i, [i, i, C.POINTER(C.c_double)]
        lib.cvtx_b200_last_pair_kernel_ms.restype, lib.cvtx_b200_last_pair_kernel_ms.argtypes = f, [i]
        lib.cvtx_b200_tune.restype, lib.cvtx_b200_tune.argtypes = None, [i, i]
        lib.cvtx_b200_guarded_only.restype, lib.cvtx_b200_guarded_only.argtypes = None, [i]
        lib.cvtx_b200_f3d_mode.restype, lib.cvtx_b200_f3d_mode.argtypes = None, [i]
        lib.cvtx_b200_sparse_route.restype, lib.cvtx_b200_sparse_route.argtypes = None, [i]
        lib.cvtx_b200_last_dispatch.restype = i
        lib.cvtx_b200_last_devices_used.restype = i
        lib.cvtx_b200_last_error.restype = C.c_char_p

    # ---- devices ----
    def device_count(self) -> int:
        return int(self.lib.cvtx_b200_device_count())

    def device_name(self, d: int):
        s = self.lib.cvtx_b200_device_name(d)
        return None if s is None else s.decode()

    def sm_count(self, d: int) -> int:
        return int(self.lib.cvtx_b200_device_sm_count(d))

    def clock_khz(self, d: int) -> int:
        return int(self.lib.cvtx_b200_device_clock_khz(d))

    def require_gpu(self) -> int:
        n = self.device_count()
        if n <= 0:
            raise BackendError("no CUDA device visible to libcvortex.so (%s); the all-pairs path has no CPU "
                               "fallback" % (self.last_error() or "cudaGetDeviceCount returned 0"))
        return n

    # ---- metadata ----
    def op_info(self, op: str, reg: str = "winckelmans") -> dict:
        v = [C.c_int() for _ in range(5)]
        rc = self.lib.cvtx_b200_op_info(OPS[op], REGS[reg], *[C.byref(x) for x in v])
        if rc:
            raise BackendError(f"{op}/{reg}: {self.last_error()}")
        keys = ("src_cols", "tgt_cols", "out_cols", "lane_ops", "sfu_ops")
        return dict(zip(keys, (x.value for x in v)))

    def plan(self, op: str, device: int, n_src: int, n_tgt: int) -> dict:
        v = [C.c_int() for _ in range(4)]
        rc = self.lib.cvtx_b200_plan(OPS[op], device, n_src, n_tgt, *[C.byref(x) for x in v])
        if rc:
            raise BackendError(self.last_error())
        return dict(zip(("block", "tgt_per_thread", "grid_x", "grid_y"), (x.value for x in v)))

    def measure_peak(self, device: int, what: str) -> float:
        """Measured pipe peak of `device`: "fp32" -> FP32 lane-ops/s (FFMA2 loop), "mufu" -> MUFU ops/s."""
        v = C.c_double()
        rc = self.lib.cvtx_b200_measure_peak(device, {"fp32": 0, "mufu": 1}[what], C.byref(v))
        if rc:
            raise BackendError(f"cvtx_b200_measure_peak failed ({rc}): {self.last_error()}")
        return float(v.value)

    def kernel_launches(self) -> int:
        return int(self.lib.cvtx_b200_kernel_launches())

    def last_pair_kernel_ms(self, device: int) -> float:
        return float(self.lib.cvtx_b200_last_pair_kernel_ms(device))

    def tune(self, tgt_per_thread: int = 0, chunks: int = 0) -> None:
        self.lib.cvtx_b200_tune(tgt_per_thread, chunks)

    def guarded_only(self, mode: int) -> None:
        """Tests / experiments (cvtx_b200_guarded_only): 1 = every chain in the guarded pair form,
        2 = the optimistic form at any size, 0 = the default (optimistic from 16 source tiles up)."""
        self.lib.cvtx_b200_guarded_only(int(mode))

    def f3d_mode(self, mode: int) -> None:
        """Tests / experiments (cvtx_b200_f3d_mode): 0 = cancellation-free filament form, 1 = the
        form that selects per pair (long / few filaments), -1 = chosen per call from the filaments (the default)."""
        self.lib.cvtx_b200_f3d_mode(int(mode))

    def sparse_route(self, on: bool) -> None:
        """Tests / experiments (cvtx_b200_sparse_route): the tile-skipping route of cvtx_P3D_M2M_vort on or off."""
        self.lib.cvtx_b200_sparse_route(1 if on else 0)

    def last_dispatch(self) -> int:
        return int(self.lib.cvtx_b200_last_dispatch())

    def last_devices_used(self) -> int:
        return int(self.lib.cvtx_b200_last_devices_used())

    def last_error(self) -> str:
        return (self.lib.cvtx_b200_last_error() or b"").decode()

    def release(self) -> None:
        self.lib.cvtx_b200_release()

    # ---- the hot path ----
    def m2m(self, op: str, reg: str, device: int, stream, src, n_src: int, tgt, n_tgt: int, out,
            sigma: float = 1.0, nu: float = 0.0) -> None:
        """Asynchronous all-pairs call on device pointers (see cvtx_b200_m2m)."""
        rc = self.lib.cvtx_b200_m2m(OPS[op], REGS[reg], device, _ptr(stream), _ptr(src), n_src,
                                    _ptr(tgt), n_tgt, _ptr(out), sigma, nu)
        if rc:
            raise BackendError(f"cvtx_b200_m2m({op}, {reg}) failed ({rc}): {self.last_error()}")

    def m2m_sharded(self, op: str, reg: str, devices, src_shards, n_src_shard, tgts, n_tgt, outs,
                    sigma: float = 1.0, nu: float = 0.0) -> None:
        """Several devices from this one process, sources sharded (see cvtx_b200_m2m_sharded): lists of
        per-device source shards, targets and outputs (device pointers / tensors on devices[g]).  The shards
        are all-gathered over NCCL inside the library; returns when every device has finished."""
        G = len(devices)
        ia = lambda v: (C.c_int * G)(*[int(x) for x in v])
        pa = lambda v: (C.c_void_p * G)(*[_ptr(x) for x in v])
        rc = self.lib.cvtx_b200_m2m_sharded(OPS[op], REGS[reg], G, ia(devices), pa(src_shards), ia(n_src_shard),
                                            pa(tgts), ia(n_tgt), pa(outs), sigma, nu)
        if rc:
            raise BackendError(f"cvtx_b200_m2m_sharded({op}, {reg}) failed ({rc}): {self.last_error()}")

    def exchange_backend(self) -> str:
        """How source shards travel between devices in this process ("nccl 2.x.y, ..." / "peer-to-peer copies ...")."""
        return (self.lib.cvtx_b200_exchange_backend() or b"").decode()

    def f3d_inf_mtrx(self, device: int, stream, fil, n_fil: int, mes, dirs, n_mes: int, out) -> None:
        """Asynchronous dense influence matrix on device pointers (see cvtx_b200_f3d_inf_mtrx)."""
        rc = self.lib.cvtx_b200_f3d_inf_mtrx(device, _ptr(stream), _ptr(fil), n_fil, _ptr(mes), _ptr(dirs), n_mes, _ptr(out))
        if rc:
            raise BackendError(f"cvtx_b200_f3d_inf_mtrx failed ({rc}): {self.last_error()}")

    def redistribute(self, dim: int, redist: str, device: int, stream, rows, n: int, grid_density: float,
                     negligible_vort: float = 0.0, out=None, max_out: int = 0) -> int:
        """Redistribution onto a grid on device pointers (see cvtx_b200_redistribute): `rows` holds n
        cvtx_P3D / cvtx_P2D structs on `device`, `out` has room for max_out (None: count only).
        Returns the number of particles created."""
        n_out = C.c_int(0)
        rc = self.lib.cvtx_b200_redistribute(dim, REDISTS[redist], device, _ptr(stream), _ptr(rows), n, grid_density,
                                             negligible_vort, _ptr(out) if out is not None else None,
                                             max_out if out is not None else 0, C.byref(n_out))
        if rc:
            raise BackendError(f"cvtx_b200_redistribute({dim}D, {redist}) failed ({rc}): {self.last_error()}")
        return n_out.value

    def pedrizzetti_relaxation(self, reg: str, device: int, stream, rows, n: int, fdt: float, sigma: float) -> None:
        """cvtx_P3D_pedrizzetti_relaxation on n cvtx_P3D structs resident on `device`, in place
        (see cvtx_b200_pedrizzetti_relaxation).  Asynchronous on `stream`."""
        rc = self.lib.cvtx_b200_pedrizzetti_relaxation(REGS[reg], device, _ptr(stream), _ptr(rows), n, fdt, sigma)
        if rc:
            raise BackendError(f"cvtx_b200_pedrizzetti_relaxation({reg}) failed ({rc}): {self.last_error()}")

    def m2m_host(self, op: str, reg: str, device: int, src: np.ndarray, tgt: np.ndarray,
                 sigma: float = 1.0, nu: float = 0.0, out: np.ndarray | None = None):
        """Synchronous call on numpy arrays; returns (result, h2d_bytes, d2h_bytes)."""
        info = self.op_info(op, reg)
        src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, info["src_cols"])
        tgt = np.ascontiguousarray(tgt, dtype=np.float32).reshape(-1, info["tgt_cols"])
        if out is None:
            out = np.empty((tgt.shape[0], info["out_cols"]), dtype=np.float32)
        up, down = C.c_size_t(), C.c_size_t()
        rc = self.lib.cvtx_b200_m2m_host(OPS[op], REGS[reg], device, src.ctypes.data, src.shape[0],
                                         tgt.ctypes.data, tgt.shape[0], out.ctypes.data, sigma, nu,
                                         C.byref(up), C.byref(down))
        if rc:
            raise BackendError(f"cvtx_b200_m2m_host({op}, {reg}) failed ({rc}): {self.last_error()}")
        return out, up.value, down.value
