"""Generate tests/golden/reference_remesh.npz from the reference's own implementation.

Run in the build container (needs /root/reference -> oracle/_ref/libcvortex_ref.so):
    python tests/golden/make_golden_remesh.py
cvtx_P3D_redistribute_on_grid, cvtx_P2D_redistribute_on_grid (five interpolants, with and
without pruning, with a too-small output array) and cvtx_P3D_pedrizzetti_relaxation are run by
the UNMODIFIED reference on small seeded inputs; inputs and outputs are stored so that the GPU
box, which has no /root/reference, can check the oracle port and the CUDA path against the
reference's actual numbers.  (The reference's two grid-tree units are compiled through
oracle/msvc_shim/ref_tree_msvc.h: their g++ branch is an `assert(false)` stub.)
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cvortex_b200.abi import CvtxLibrary  # noqa: E402
from oracle import binding  # noqa: E402
from util import remesh_cases, remesh_particles  # noqa: E402

N = 300
RELAX = (("winckelmans", 0.3, 0.1), ("gaussian", 0.15, 0.25), ("planetary", 0.5, 0.05))


def main():
    binding.build(ref=True)
    assert binding.have_ref(), "oracle/_ref/libcvortex_ref.so missing (needs /root/reference)"
    ref = CvtxLibrary(binding.REF_SO)
    ref.initialise()
    out = {}
    for dim, name, h, negl, cap in remesh_cases():
        key = f"remesh|{dim}|{name}|{negl}|{cap}"
        rng = np.random.default_rng(zlib.crc32(key.encode()))
        p = remesh_particles(rng, N, dim)
        fn = ref.P3D_redistribute_on_grid if dim == 3 else ref.P2D_redistribute_on_grid
        out[key + "|in"] = p
        out[key + "|out"] = fn(p, name, h, negl, max_output=cap)
        out[key + "|count"] = np.array([fn(p, name, h, negl, count_only=True)])
    for reg, sigma, fdt in RELAX:
        key = f"relax|{reg}"
        rng = np.random.default_rng(zlib.crc32(key.encode()))
        p = remesh_particles(rng, N, 3)
        out[key + "|in"] = p
        out[key + "|par"] = np.array([sigma, fdt], np.float32)
        out[key + "|out"] = ref.P3D_pedrizzetti_relaxation(p, fdt, reg, sigma)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_remesh.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
