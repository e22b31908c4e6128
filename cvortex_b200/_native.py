"""Locate, (re)build and load ``libcvortex.so`` -- the CUDA library IS the product.

There is no Python or CPU stand-in for it: if the shared object is missing or
cannot be loaded, importing the API raises.  ``build()`` drives
``cvortex_b200/csrc/Makefile`` (nvcc, ``-gencode arch=compute_100a,code=sm_100a``);
the result lives in-tree at ``cvortex_b200/lib/libcvortex.so``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "lib", "libcvortex.so")


class NativeLibraryError(RuntimeError):
    pass


def build(jobs: int = 4, verbose: bool = False) -> str:
    """Compile every CUDA / C++ source of the library for sm_100a (no GPU needed)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs}", "--no-print-directory"]
    res = subprocess.run(cmd, stdout=None if verbose else subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise NativeLibraryError("building libcvortex.so failed:\n" + (res.stdout or ""))
    return LIB_PATH


def preload_python_nccl() -> None:
    """libcvortex.so opens ``libnccl.so.2`` at run time, when a call first spans several devices.  In a Python
    process that will also import torch, the NCCL that must win is the one torch was built against (the
    ``nvidia-nccl`` wheel next to it): the loader keeps ONE library per soname, and a system NCCL loaded first
    leaves torch with unresolved symbols (seen: ``ncclDevCommCreate``).  Loading the wheel's copy here, before
    anything else asks for the soname, makes both use it.  No wheel -> nothing to do: C hosts get the system one."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    except (ImportError, ValueError):
        roots = []
    for root in roots:
        path = os.path.join(root, "lib", "libnccl.so.2")
        if os.path.exists(path):
            try:
                C.CDLL(path, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load() -> C.CDLL:
    preload_python_nccl()
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cvortex_b200/csrc`). There is no fallback implementation.")
    try:
        return C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
