"""Generate tests/golden/reference_m2m.npz from the reference's own CPU path.

Run in the build container (needs /root/reference -> oracle/_ref/libcvortex_ref.so):
    python tests/golden/make_golden.py
Every M2M op x regularisation the reference accelerates is evaluated by the UNMODIFIED
reference (compiled by oracle/Makefile) on small seeded inputs in three regimes; inputs and
outputs are stored so that the GPU box, which has no /root/reference, can still check the
oracle port and the CUDA path against the reference's actual numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cvortex_b200.abi import CvtxLibrary  # noqa: E402
from oracle import binding  # noqa: E402
from util import call_abi, make_case, op_cases  # noqa: E402

REGIMES = {"overlap": (10.0, 0.3, 0.1), "bench": (10.0, 0.02, 1.0), "tiny": (1.53e-4, 0.3, 0.1)}
N, M = 192, 96


def main():
    binding.build(ref=True)
    assert binding.have_ref(), "oracle/_ref/libcvortex_ref.so missing (needs /root/reference)"
    ref = CvtxLibrary(binding.REF_SO)
    ref.initialise()
    out = {}
    for regime, (box, sigma, nu) in REGIMES.items():
        for k, (op, reg) in enumerate(op_cases() + [("P3D_M2M_vort", r) for r in ("winckelmans", "planetary", "gaussian")]):
            rng = np.random.default_rng(abs(hash_name(regime, op, reg)))
            if op == "P3D_M2M_vort":
                src, tgt = make_case("P3D_M2M_vel", rng, N, M, box=box, self_targets=True)
            else:
                src, tgt = make_case(op, rng, N, M, box=box, self_targets=True)
            res = call_abi(ref, op, src, tgt, reg, sigma, nu)
            key = f"{regime}|{op}|{reg}"
            out[key + "|src"], out[key + "|tgt"], out[key + "|out"] = src, tgt, np.asarray(res, np.float32)
            out[key + "|par"] = np.array([sigma, nu], np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_m2m.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out) // 4, "cases")


def hash_name(*names):
    import zlib
    return zlib.crc32("/".join(names).encode())


if __name__ == "__main__":
    main()
