"""cvortex_b200 -- B200-native backend for cvortex's all-pairs vortex interactions.

The product is ``cvortex_b200/lib/libcvortex.so`` (hand-written sm_100a CUDA
kernels behind the reference's unchanged ``cvtx_*`` C ABI, plus the thin
``cvtx_b200_*`` device-level ABI).  This package is the host-side mirror used
by the tests and the benchmark:

* :mod:`cvortex_b200.api`      reference-named calls through the ``cvtx_*`` ABI
* :mod:`cvortex_b200.abi`      the ctypes view of that ABI (any conforming library)
* :mod:`cvortex_b200.device`   device-pointer level (``cvtx_b200_m2m``)
* :mod:`cvortex_b200.sharding` target partitioning / source all-gather for one
  process per GPU under ``torch.distributed``
"""
from . import _native  # noqa: F401

__all__ = ["api", "abi", "device", "sharding"]
__version__ = "0.3.8+b200.1"

# In a Python process the NCCL that libcvortex.so opens at run time must be the one torch uses (see _native).
_native.preload_python_nccl()
