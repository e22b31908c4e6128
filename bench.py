#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 all-pairs backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): billion pair-interactions per second.  Default workload =
BASELINE.json configs[1]: cvtx_P3D_M2M_vel + cvtx_P3D_M2M_dvort, 1M particles, Gaussian
regularisation; a STEP is one pass of both ops (2 x 1e12 pair-interactions).  Inputs follow
the reference's benchmark (bench/bencharraysetup.c:43-58, bench/benchP3D.c:283-415): coords and
vorticity uniform in [0,10), volume 0.01, sigma 0.02; vel targets are an independent uniform
point cloud, dvort targets are the particles themselves.  Seeds are fixed.

  value   whole-job throughput with the particles resident in HBM (sharded over the ranks;
          each step all-gathers the source rows over NCCL when N > 1), CUDA-event timed.
  e2e     the same work through the reference's own ABI -- cvtx_P3D_M2M_vel / _dvort of
          libcvortex.so with HOST arrays of pointers: gather, H2D, kernels, D2H inside the
          timed region (wall clock, max over ranks).
  roofline  the dominant kernel's FP32 lane-op rate against the chip's FP32 issue peak.
  cpu_baseline  the reference's own OpenMP CPU path (oracle/_ref) on this box's cores, on a
          bounded target sample of the same workload.

With N > 1 the targets are sharded over the ranks (total work fixed: "strong" scaling).
`--impl reference` times the unmodified reference CPU path instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIGMA, NU = 0.02, 1.0
WORKLOADS = {
    # name: (n sources, n targets, [(op, reg)])      -- BASELINE.json configs[0..4] + the north-star headline
    "p3d_vel_winckelmans_10k": (10_000, 10_000, [("P3D_M2M_vel", "winckelmans")]),                                   # configs[0]
    "p3d_vel+dvort_gaussian_1M": (1_000_000, 1_000_000, [("P3D_M2M_vel", "gaussian"), ("P3D_M2M_dvort", "gaussian")]),  # configs[1]
    "p3d_visc_winckelmans_4M": (4_000_000, 4_000_000, [("P3D_M2M_visc_dvort", "winckelmans")]),                      # configs[2]
    "p2d_vel+visc_gaussian_4M": (4_000_000, 4_000_000, [("P2D_M2M_vel", "gaussian"), ("P2D_M2M_visc_dvort", "gaussian")]),  # configs[3]
    "f3d_vel+dvort_100k_on_2M": (100_000, 2_000_000, [("F3D_M2M_vel", "singular"), ("F3D_M2M_dvort", "singular")]),  # configs[4]
    "p3d_vel_winckelmans_1M": (1_000_000, 1_000_000, [("P3D_M2M_vel", "winckelmans")]),                              # headline target
}
DEFAULT_WORKLOAD = "p3d_vel+dvort_gaussian_1M"
PARTICLE_TARGETS = {"P3D_M2M_dvort", "P3D_M2M_visc_dvort", "P2D_M2M_visc_dvort", "F3D_M2M_dvort"}


def make_inputs(n, m, ops):
    """Seeded version of the reference benchmark's arrays (one stream per array).
    Returns (S sources, TP particle targets, TX point targets).  For the particle ops the
    particle targets ARE the sources (self-interaction, as bench/benchP3D.c:283-415 does) and
    the point targets an independent uniform cloud; filaments (no upstream bench) follow
    SURVEY 8d: start uniform, end = start + U(-0.1, 0.1)^3, strength U[0, 10), acting on m
    independent particles / points."""
    two_d, fil = ops[0][0].startswith("P2D"), ops[0][0].startswith("F3D")
    rng_p, rng_x, rng_t = (np.random.default_rng(s) for s in (20261017, 20261018, 20261019))
    if two_d:
        S = rng_p.uniform(0.0, 10.0, (n, 4)).astype(np.float32)
        S[:, 3] = 0.01
        TX = rng_x.uniform(0.0, 10.0, (m, 2)).astype(np.float32)
        TP = S[:m] if m <= n else None
    elif fil:
        S = rng_p.uniform(0.0, 10.0, (n, 7)).astype(np.float32)
        S[:, 3:6] = S[:, 0:3] + rng_x.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
        TP = rng_t.uniform(0.0, 10.0, (m, 7)).astype(np.float32)
        TP[:, 6] = 0.01
        TX = np.ascontiguousarray(TP[:, :3])
    else:
        S = rng_p.uniform(0.0, 10.0, (n, 7)).astype(np.float32)
        S[:, 6] = 0.01
        TX = rng_x.uniform(0.0, 10.0, (m, 3)).astype(np.float32)
        TP = S[:m] if m <= n else None
    return S, TP, TX


def out_cols(op):
    return {"P2D_M2M_vel": 2, "P2D_M2M_visc_dvort": 1}.get(op, 3)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference arm / cpu baseline
def time_reference_cpu(n, m, ops, n_sample_targets, repeats=1):
    """Time the reference's own OpenMP CPU path (oracle/_ref, else the oracle port) on all n
    sources x a strided sample of targets.  Returns (Gpair/s, seconds, kind, cores, description)."""
    from oracle import binding
    P, TP, X = make_inputs(n, m, ops)
    idx = np.arange(0, m, max(1, m // n_sample_targets))[:n_sample_targets]
    Xs, Ps = np.ascontiguousarray(X[idx]), np.ascontiguousarray(TP[idx])
    ora = binding.Oracle()
    cores = ora.num_threads()
    if binding.have_ref():
        from cvortex_b200.abi import CvtxLibrary
        ref = CvtxLibrary(binding.REF_SO)
        ref.initialise()
        kind = "reference"

        def run(op, reg):
            fn = getattr(ref, op)
            tg = Ps if op in PARTICLE_TARGETS else Xs
            if op.startswith("F3D"):
                return fn(P, tg)
            return fn(P, tg, reg, SIGMA, NU) if op.endswith("visc_dvort") else fn(P, tg, reg, SIGMA)
    else:
        kind = "port"

        def run(op, reg):
            return ora.m2m(op, P, Ps if op in PARTICLE_TARGETS else Xs, reg, SIGMA, NU)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for op, reg in ops:
            run(op, reg)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    pairs = float(n) * len(idx) * len(ops)
    desc = (f"all {n} sources x {len(idx)} stride-sampled targets, {'+'.join(o for o, _ in ops)}, "
            f"{pairs:.2e} pair-interactions per pass")
    return pairs / best / 1e9, best, kind, cores, desc


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n, m, ops = WORKLOADS[args.workload]
    from oracle import binding
    cores = binding.Oracle().num_threads()
    # bounded sample: ~2-4 s per step on this box's cores
    m_s = max(64, min(m, 32 * cores)) if n >= 100_000 else m
    times = []
    for k in range(args.warmup + args.steps):
        rate, dt, kind, cores, desc = time_reference_cpu(n, m, ops, m_s)
        if k >= args.warmup:
            times.append(dt)
    pairs = float(n) * min(m_s, m) * len(ops)
    total = sum(times)
    value = pairs * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "pair-interactions/s", "value": value, "unit": "Gpair/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n, m, ops, world=args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gpair/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "Gpair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, n, m, ops, world):
    return {"workload": name, "ops": [f"cvtx_{o}/{r}" for o, r in ops], "n_sources": n, "n_targets": m,
            "sigma": SIGMA, "kinematic_visc": NU, "pair_interactions_per_step": float(n) * m * len(ops),
            "parallelism": f"targets sharded over {world} GPU(s), sources replicated",
            "l2": "256 MiB scratch write between steps (L2 flush); sources (32 MB packed) are L2-resident by design"}


# ------------------------------------------------------------------ the B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cvortex_b200 import api
    from cvortex_b200.sharding import ShardedM2M, target_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the all-pairs path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    api.initialise(require_gpu=True)
    api.use_only(local_rank)
    be = api.backend()
    lib = api.library()

    n, m, ops = WORKLOADS[args.workload]
    P, TP, X = make_inputs(n, m, ops)
    lo, hi = target_range(m, rank, world)
    m_local = hi - lo
    slo, shi = target_range(n, rank, world)
    # HBM-resident state: this rank's shard of the sources and its contiguous range of targets
    src_local = torch.from_numpy(np.ascontiguousarray(P[slo:shi])).to(dev)
    tgts = {op: torch.from_numpy(np.ascontiguousarray((TP if op in PARTICLE_TARGETS else X)[lo:hi])).to(dev) for op, _ in ops}
    outs = {op: torch.empty((m_local, out_cols(op)), device=dev) for op, _ in ops}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sharded = ShardedM2M(be, local_rank, n)
    stream = torch.cuda.current_stream()

    ev = {op: [] for op, _ in ops}

    def step(record):
        full = sharded.gather_sources(src_local)            # NCCL all-gather of the raw rows when world > 1
        for op, reg in ops:
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
            sharded.step(op, reg, src_local, tgts[op], outs[op], SIGMA, NU, stream=stream.cuda_stream, src_full=full)
            if record:
                b.record(stream)
                ev[op].append((a, b))
        flush.zero_()                                        # evict L2 between steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = be.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = be.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    nl = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    pairs_per_step = float(n) * m * len(ops)
    value = pairs_per_step * args.steps / (ms * 1e-3) / 1e9

    # per-kernel durations (events on the launching stream, inside the timed region)
    kern = {}
    for op, reg in ops:
        d = np.array([a.elapsed_time(b) for a, b in ev[op]])
        info = be.op_info(op, reg)
        kern[op] = {"reg": reg, "ms": float(d.mean()), "lane_ops": info["lane_ops"], "sfu_ops": info["sfu_ops"],
                    "pairs": float(n) * m_local}

    # ---- config 2 "as fused pass" (SURVEY 8d): velocity AT the particles + stretching in one
    # sweep over the sources (thin-ABI op CVTX_B200_P3D_VEL_DVORT); counted as 2 pair-interactions
    # per (source, target) like the two separate ops it replaces
    fused = None
    if [o for o, _ in ops] == ["P3D_M2M_vel", "P3D_M2M_dvort"]:
        reg = ops[0][1]
        out6 = torch.empty((m_local, 6), device=dev)
        full = sharded.gather_sources(src_local)
        for _ in range(2):
            be.m2m("P3D_M2M_vel_dvort", reg, local_rank, stream.cuda_stream, full, n, tgts["P3D_M2M_dvort"], m_local, out6, SIGMA, NU)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            full = sharded.gather_sources(src_local)
            be.m2m("P3D_M2M_vel_dvort", reg, local_rank, stream.cuda_stream, full, n, tgts["P3D_M2M_dvort"], m_local, out6, SIGMA, NU)
            flush.zero_()
        f1.record(stream)
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fms = float(tf.item()) / args.steps
        finfo = be.op_info("P3D_M2M_vel_dvort", reg)
        fused = {"value": 2.0 * n * m / (fms * 1e-3) / 1e9, "unit": "Gpair/s", "ms_per_step": fms,
                 "lane_ops_per_source_target": finfo["lane_ops"],
                 "note": "vel evaluated at the particle positions (not the independent point cloud of the "
                         "separate-op step) fused with dvort; additive thin-ABI op, not part of `value`"}

    # ---- e2e: the reference's ABI with host pointer arrays, wall clock
    e2e = None
    if not args.no_e2e:
        # what a C caller holds between calls: the struct arrays, the arrays of pointers into them
        # (built once, as the reference's bench setup does) and preallocated result arrays
        from cvortex_b200.abi import PointerRows
        srcs = PointerRows(P, P.shape[1])
        host_t, host_o = {}, {}
        for op, _ in ops:
            rows = np.ascontiguousarray((TP if op in PARTICLE_TARGETS else X)[lo:hi])
            host_t[op] = PointerRows(rows, rows.shape[1]) if op in PARTICLE_TARGETS else rows
            host_o[op] = np.empty((m_local, out_cols(op)), dtype=np.float32)

        def e2e_step():
            res = []
            for op, reg in ops:
                fn = getattr(lib, op)
                if op.startswith("F3D"):
                    res.append(fn(srcs, host_t[op], out=host_o[op]))
                elif op.endswith("visc_dvort"):
                    res.append(fn(srcs, host_t[op], reg, SIGMA, NU, out=host_o[op]))
                else:
                    res.append(fn(srcs, host_t[op], reg, SIGMA, out=host_o[op]))
                assert be.last_dispatch() == 1
            return res
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(P.nbytes + (host_t[op].rows if op in PARTICLE_TARGETS else host_t[op]).nbytes for op, _ in ops)
        d2h = sum(r.nbytes for r in res)
        e2e = {"value": pairs_per_step * args.steps / dt / 1e9, "unit": "Gpair/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt / args.steps,
               "api": "cvtx_*_M2M_* C ABI, host arrays of pointers; per rank when N > 1"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        sms = be.sm_count(local_rank)
        sm_max_mhz = float(peaks.get("sm_max_mhz", be.clock_khz(local_rank) / 1e3))
        peak_lane = sms * 128 * sm_max_mhz * 1e6                   # FP32 lane-ops/s (FMA = 1 lane-op)
        peak_sfu = sms * 16 * sm_max_mhz * 1e6
        dom = max(kern, key=lambda k: kern[k]["ms"])
        kd = kern[dom]
        rate = kd["pairs"] / (kd["ms"] * 1e-3)
        bound = "sfu" if kd["sfu_ops"] * 8 > kd["lane_ops"] else "fp32"
        if bound == "fp32":
            achieved, peak, unit = rate * kd["lane_ops"] * 2 / 1e12, peak_lane * 2 / 1e12, "TFLOP/s"
        else:
            achieved, peak, unit = rate * kd["sfu_ops"] / 1e12, peak_sfu / 1e12, "T MUFU-op/s"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{dom}/{kd['reg']}")
        roofline = {
            "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak, "traffic": traffic,
            "kernel": f"m2m_kernel<{dom}/{kd['reg']}>", "avg_launch_ms": kd["ms"], "pairs_per_launch": kd["pairs"],
            "lane_ops_per_pair": kd["lane_ops"], "sfu_ops_per_pair": kd["sfu_ops"],
            "traffic_note": ("DRAM bytes per launch (ncu); the inputs are ~70 MB, the rest are the FP64 partial sums of the "
                             "source chunks that exist for load balance (DESIGN.md section 3) -- 0.1 ms of HBM time in a "
                             "1.2 s FP32-pipe-bound launch"),
            "peak_source": (f"nominal FP32 issue peak = {sms} SMs x 128 lanes x {sm_max_mhz:.0f} MHz (sm_max_mhz of "
                            "MEASURED_PEAKS.json) x 2 flop; each algorithmic FP32 lane-op counted as one FMA slot. "
                            "MEASURED_PEAKS.json has no FP32 figure (HBM and bf16 only); an FFMA-only micro-benchmark "
                            "sustains 97.4% of this number on this pool (profiles/ubench_r1.txt)"),
            "all_kernels": {f"{k}/{v['reg']}": {"avg_launch_ms": v["ms"], "gpairs_per_s": v["pairs"] / (v["ms"] * 1e-3) / 1e9,
                                               "frac_fp32": v["pairs"] / (v["ms"] * 1e-3) * v["lane_ops"] / peak_lane,
                                               "frac_sfu": v["pairs"] / (v["ms"] * 1e-3) * v["sfu_ops"] / peak_sfu}
                            for k, v in kern.items()},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding
            cores = binding.Oracle().num_threads()
            m_s = max(256, min(m, 96 * cores)) if n >= 100_000 else m
            rate_c, secs, kind, cores, desc = time_reference_cpu(n, m, ops, m_s)
            cpu = {"value": rate_c, "unit": "Gpair/s", "cores": cores, "kind": kind, "sample": desc + f", {secs:.1f} s"}
        line = {
            "metric": "pair-interactions/s", "value": value, "unit": "Gpair/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, n, m, ops, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(nl.item()),
            "roofline": roofline, "cpu_baseline": cpu, "fused_vel_dvort": fused,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
