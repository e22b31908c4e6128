"""Redistribution onto a grid: the B200 path beside the reference's own CPU implementation.

    python tools/remesh_bench.py [--sizes 10000,80000,250000,1000000] [--no-cpu]

Workload = the reference's benchmark (bench/benchredistribution.c:49-75): n particles uniform in
the unit box, grid spacing cbrt(2/n) (about two particles per cell in 3-D), negligible_vort 1e-4,
room for 4n created particles; every interpolant, 3-D and 2-D.  Each line reports the wall time
of the whole public-ABI call (host pointer arrays in, host particle array out) for
  gpu : libcvortex.so with accelerator 0 enabled (remesh_device.cu + the host pruning stage),
  cpu : the reference's implementation compiled from /root/reference (oracle/_ref, all host cores)
        -- test infrastructure, timed here only as the baseline,
and checks the two results against each other (same nodes, same order, strengths to 2e-6).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cvortex_b200 import _native  # noqa: E402
from cvortex_b200.abi import CvtxLibrary, PointerRows  # noqa: E402
from util import REDISTS, assert_same_remesh, remesh_particles  # noqa: E402


def timed(fn, repeats):
    best = float("inf")
    out = None
    for _ in range(repeats):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10000,80000,250000,1000000")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-max", type=int, default=1000000, help="largest n the CPU reference is timed on")
    args = ap.parse_args()
    lib = CvtxLibrary(_native.LIB_PATH)
    lib.initialise()
    assert lib.num_enabled_accelerators() > 0, "no CUDA accelerator enabled"
    ref = None
    if not args.no_cpu:
        from oracle import binding
        if binding.have_ref():
            ref = CvtxLibrary(binding.REF_SO)
            ref.initialise()
    print(f"accelerator: {lib.accelerator_name(0)}; host cores: {os.cpu_count()}; cpu reference: {'oracle/_ref' if ref else 'not timed'}")
    for dim in (3, 2):
        for n in (int(s) for s in args.sizes.split(",")):
            p = PointerRows(remesh_particles(np.random.default_rng(n + dim), n, dim, signed=False), 7 if dim == 3 else 4)
            h = float(np.cbrt(2.0 / n))
            # the caller's output array, allocated once as the reference benchmark does
            # (bench/benchredistribution.c:50: create_particles_3D_outarr)
            out_gpu = np.zeros((4 * n, 7 if dim == 3 else 4), np.float32)
            out_cpu = np.zeros_like(out_gpu)
            for name in REDISTS:
                fn = lib.P3D_redistribute_on_grid if dim == 3 else lib.P2D_redistribute_on_grid
                fn(p, name, h, 1e-4, max_output=4 * n, out=out_gpu)   # warm-up: buffers grow once
                t_gpu, got = timed(lambda: fn(p, name, h, 1e-4, max_output=4 * n, out=out_gpu), 3)
                line = f"{dim}D {name:8s} n={n:8d} -> {len(got):8d} particles | gpu {t_gpu * 1e3:9.2f} ms  {n / t_gpu / 1e6:8.2f} Mparticle/s"
                if ref is not None and n <= args.cpu_max:
                    rfn = ref.P3D_redistribute_on_grid if dim == 3 else ref.P2D_redistribute_on_grid
                    t_cpu, want = timed(lambda: rfn(p, name, h, 1e-4, max_output=4 * n, out=out_cpu), 1)
                    assert_same_remesh(got, want, what=f"{dim}D {name} {n}")
                    line += f" | cpu {t_cpu * 1e3:10.2f} ms  {n / t_cpu / 1e6:7.3f} Mparticle/s | x{t_cpu / t_gpu:7.1f}  parity ok"
                print(line, flush=True)


if __name__ == "__main__":
    main()
