"""A vortex-particle time loop that never leaves the GPU (the additive device-pointer API).

    python examples/vortex_ring.py [--particles 20000] [--steps 20] [--remesh-every 5]

A thick-cored vortex ring is discretised into particles (cvtx_P3D rows in one torch tensor).
Every step is what a cvortex user's loop does, with each library call replaced by its
device-pointer twin of include/cvtx_b200.h:

    u       = cvtx_P3D_M2M_vel   (particles -> their own positions)     cvtx_b200_m2m
    d alpha = cvtx_P3D_M2M_dvort (vortex stretching)                    cvtx_b200_m2m
              (by default both in ONE pass over the pairs: the fused CVTX_B200_P3D_VEL_DVORT op)
    x += u dt;  alpha += d alpha dt                                      (torch, in place)
    every k steps: cvtx_P3D_redistribute_on_grid (M4')                   cvtx_b200_redistribute
                   cvtx_P3D_pedrizzetti_relaxation (optional)            cvtx_b200_pedrizzetti_relaxation

No particle data crosses PCIe; only the particle count comes back to the host after a
redistribution.  The script prints the ring's position, the invariants of the motion (total
vorticity, linear impulse) and the time per step; `run()` is also what
tests/test_gpu_timestep.py drives.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def ring_particles(n_target: int, radius: float = 1.0, core: float = 0.2, circulation: float = 1.0):
    """Particles on a regular lattice inside the torus (R = radius, core radius = core), each
    carrying alpha = omega * volume with a Gaussian vorticity profile across the core.
    Returns (rows (n, 7) float32, lattice spacing)."""
    volume = 2.0 * np.pi**2 * radius * core**2
    h = float((volume / n_target) ** (1.0 / 3.0))
    m = int(np.ceil((radius + core) / h)) + 1
    ax = (np.arange(-m, m + 1) + 0.5) * h
    x, y, z = np.meshgrid(ax, ax, ax[np.abs(ax) <= core + h], indexing="ij")
    x, y, z = x.ravel(), y.ravel(), z.ravel()
    rho = np.hypot(x, y)
    d2 = (rho - radius) ** 2 + z**2
    keep = d2 <= core**2
    x, y, z, rho, d2 = x[keep], y[keep], z[keep], rho[keep], d2[keep]
    s = core / 2.0
    omega = circulation / (2.0 * np.pi * s**2) * np.exp(-d2 / (2.0 * s**2))      # azimuthal vorticity
    rows = np.zeros((len(x), 7), np.float32)
    rows[:, 0], rows[:, 1], rows[:, 2] = x, y, z
    rows[:, 3] = -y / rho * omega * h**3
    rows[:, 4] = x / rho * omega * h**3
    rows[:, 6] = h**3
    return rows, h


def invariants(torch, rows):
    """Total vorticity (zero for a closed ring) and linear impulse 1/2 sum x cross alpha."""
    x, a = rows[:, :3].double(), rows[:, 3:6].double()
    return a.sum(0).cpu().numpy(), 0.5 * torch.cross(x, a, dim=1).sum(0).cpu().numpy()


def run(n_particles=20000, steps=20, remesh_every=5, dt=0.05, reg="gaussian", verbose=True, fused=True, relax=0.0):
    import torch
    from cvortex_b200 import api

    api.initialise(require_gpu=True)
    api.use_only(0)
    dev = api.backend()
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream

    host_rows, h = ring_particles(n_particles)
    sigma = 1.5 * h
    n, room = len(host_rows), 8 * len(host_rows)          # redistribution may create more particles than it gets
    rows, spare = torch.empty((room, 7), device="cuda"), torch.empty((room, 7), device="cuda")
    rows[:n] = torch.from_numpy(host_rows).cuda()
    vel = torch.empty((room, 3), device="cuda")
    dalpha = torch.empty_like(vel)
    both = torch.empty((room, 6), device="cuda")          # fused op: u and d alpha side by side
    history = []
    tick, tock = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for step in range(steps + 1):
        total, impulse = invariants(torch, rows[:n])
        z = float((rows[:n, 2].double() * rows[:n, 3:6].double().norm(dim=1)).sum() / rows[:n, 3:6].double().norm(dim=1).sum())
        history.append({"step": step, "n": n, "z": z, "total": total, "impulse": impulse})
        if verbose:
            print(f"step {step:3d}  n={n:7d}  ring z={z:+.4f}  |sum alpha|={np.linalg.norm(total):.2e}  impulse_z={impulse[2]:+.5f}"
                  + (f"  {history[-2]['ms']:.2f} ms/step" if step else ""))
        if step == steps:
            break
        tick.record()
        if fused:
            dev.m2m("P3D_M2M_vel_dvort", reg, 0, stream, rows, n, rows, n, both, sigma)
            rows[:n, :6] += dt * both[:n]
        else:
            points = rows[:n, :3].contiguous()
            dev.m2m("P3D_M2M_vel", reg, 0, stream, rows, n, points, n, vel, sigma)
            dev.m2m("P3D_M2M_dvort", reg, 0, stream, rows, n, rows, n, dalpha, sigma)
            rows[:n, :3] += dt * vel[:n]
            rows[:n, 3:6] += dt * dalpha[:n]
        if remesh_every and (step + 1) % remesh_every == 0:
            n = dev.redistribute(3, "m4p", 0, stream, rows, n, h, 1e-3, spare, room)
            rows, spare = spare, rows
            if relax > 0.0:       # pull the particle strengths back towards the field they represent
                dev.pedrizzetti_relaxation(reg, 0, stream, rows, n, relax, sigma)
        tock.record()
        tock.synchronize()
        history[-1]["ms"] = tick.elapsed_time(tock)
    return history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--remesh-every", type=int, default=5)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--separate", action="store_true", help="two all-pairs calls per step instead of the fused op")
    ap.add_argument("--relax", type=float, default=0.0, help="Pedrizzetti relaxation factor f dt applied after each redistribution")
    a = ap.parse_args()
    run(a.particles, a.steps, a.remesh_every, a.dt, fused=not a.separate, relax=a.relax)
