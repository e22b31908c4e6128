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
  roofline  the dominant kernel's FP32 lane-op rate against the chip's FP32 issue peak
          (nominal, and as measured in this run by an FFMA2 / MUFU loop of the library).
  cpu_baseline  the reference's own OpenMP CPU path (oracle/_ref) on this box's cores, on a
          bounded target sample of the same workload (all sources x 4096 strided targets).
  parity  the GPU results on those same 4096 targets against the reference's outputs and
          against the FP64 oracle (relative L2 per output array; N = 1 only).
  inlib_multi_gpu  (N > 1) the same step from ONE process with all N accelerators enabled through the
          reference's own cvtx_accelerator_enable: sharded upload + NCCL all-gather inside the library.
  extra   the other BASELINE configs and the north-star headline (cvtx_P3D_M2M_vel,
          Winckelmans, 1M) measured the same way at reduced step counts, each with its own parity
          sample (512 stride-sampled targets against all sources; N = 1 only).

With N > 1 the targets are sharded over the ranks (total work fixed: "strong" scaling).
`--impl reference` times the unmodified reference CPU path instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes
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
CPU_SAMPLE_TARGETS = 4096            # all N sources x M' = 4096 stride-sampled targets, both CPU legs
CPU_EXTRA_SAMPLE_TARGETS = 512       # the parity sample of the `extra` configs (4M sources x 512 targets = 2e9 pairs per op on the CPU)
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
# Below this many pairs the library itself uses one device (host_api.cu kShardMinPairs); the bench does the same.
SHARD_MIN_PAIRS = 2.0e9


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


def sample_index(m, n_sample):
    """The stride sample of targets both CPU legs and the parity check use."""
    return np.arange(0, m, max(1, m // n_sample))[:n_sample]


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a))


# ------------------------------------------------------------------ host threads
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_cores():
    """Give the reference's OpenMP loops every core of the box.  Launchers such as torchrun export
    OMP_NUM_THREADS=1 to their children, which once put the reference arm on ONE core at N >= 2;
    the environment is overridden before the OpenMP runtime loads and the thread count is also set
    through the runtime's own API (the oracle libraries link the system libgomp.so.1)."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(n)
    except OSError:
        pass
    return n


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
class ReferenceCpu:
    """The reference's own OpenMP CPU path (oracle/_ref, else the oracle port) on all n sources x a
    strided sample of targets of a workload.  One object = inputs built once; pass() runs every op
    of the workload once and keeps the outputs (the parity sample)."""

    def __init__(self, n, m, ops, n_sample_targets=CPU_SAMPLE_TARGETS):
        self.cores = use_all_host_cores()
        from oracle import binding
        self.n, self.m, self.ops = n, m, ops
        self.P, TP, X = make_inputs(n, m, ops)
        self.idx = sample_index(m, n_sample_targets)
        self.Xs = np.ascontiguousarray(X[self.idx])
        self.Ps = np.ascontiguousarray(TP[self.idx]) if TP is not None else None
        self.ora = binding.Oracle()
        self.cores = min(self.cores, self.ora.num_threads()) if self.ora.num_threads() > 0 else self.cores
        if binding.have_ref():
            from cvortex_b200.abi import CvtxLibrary, PointerRows
            self.ref = CvtxLibrary(binding.REF_SO)
            self.ref.initialise()
            self.kind = "reference"
            # what a C caller holds between calls: struct arrays + pointer arrays, built once
            self.P_rows = PointerRows(self.P, self.P.shape[1])
            self.Ps_rows = PointerRows(self.Ps, self.Ps.shape[1]) if self.Ps is not None else None
        else:
            self.ref, self.kind = None, "port"
        self.pairs = float(n) * len(self.idx) * len(ops)
        self.outputs = {}

    def targets(self, op):
        return self.Ps if op in PARTICLE_TARGETS else self.Xs

    def _run(self, op, reg):
        if self.ref is None:
            return self.ora.m2m(op, self.P, self.targets(op), reg, SIGMA, NU)
        fn = getattr(self.ref, op)
        tg = self.Ps_rows if op in PARTICLE_TARGETS else self.Xs
        if op.startswith("F3D"):
            return fn(self.P_rows, tg)
        return fn(self.P_rows, tg, reg, SIGMA, NU) if op.endswith("visc_dvort") else fn(self.P_rows, tg, reg, SIGMA)

    def warm(self):
        """One short untimed pass (thread pool, page faults, caches): 1/16 of the sample."""
        keep = (self.Xs, self.Ps, getattr(self, "Ps_rows", None))
        k = max(1, len(self.idx) // 16)
        self.Xs = np.ascontiguousarray(self.Xs[:k])
        if self.Ps is not None:
            self.Ps = np.ascontiguousarray(self.Ps[:k])
            if self.ref is not None:
                from cvortex_b200.abi import PointerRows
                self.Ps_rows = PointerRows(self.Ps, self.Ps.shape[1])
        for op, reg in self.ops:
            self._run(op, reg)
        self.Xs, self.Ps = keep[0], keep[1]
        if self.ref is not None:
            self.Ps_rows = keep[2]

    def pass_(self):
        t0 = time.perf_counter()
        for op, reg in self.ops:
            self.outputs[op] = np.asarray(self._run(op, reg)).reshape(len(self.idx), -1)
        return time.perf_counter() - t0

    def f64(self, op, reg):
        return np.asarray(self.ora.m2m(op, self.P, self.targets(op), reg, SIGMA, NU, f64=True)).reshape(len(self.idx), -1)

    def describe(self, seconds=None):
        s = (f"all {self.n} sources x {len(self.idx)} stride-sampled targets, {'+'.join(o for o, _ in self.ops)}, "
             f"{self.pairs:.2e} pair-interactions per pass, {self.cores} OpenMP threads")
        return s + (f", {seconds:.1f} s" if seconds is not None else "")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n, m, ops = WORKLOADS[args.workload]
    cpu = ReferenceCpu(n, m, ops, CPU_SAMPLE_TARGETS if n >= 100_000 else m)
    cpu.warm()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu.pass_()
    times = [cpu.pass_() for _ in range(args.steps)]
    total = sum(times)
    value = cpu.pairs * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "pair-interactions/s", "value": value, "unit": "Gpair/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n, m, ops, world=args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gpair/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": cpu.describe(total / len(times)),
                         "warmup_note": "one short pass + at most one full untimed pass, whatever --warmup says: each "
                                        "pass is seconds of CPU time and nothing is left to warm after the first"},
        "e2e": {"value": value, "unit": "Gpair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(name, n, m, ops, world):
    return {"workload": name, "ops": [f"cvtx_{o}/{r}" for o, r in ops], "n_sources": n, "n_targets": m,
            "sigma": SIGMA, "kinematic_visc": NU, "pair_interactions_per_step": float(n) * m * len(ops),
            "parallelism": f"targets sharded over {world} GPU(s), sources replicated",
            "l2": "256 MiB scratch write between steps (L2 flush); sources (32 MB packed) are L2-resident by design"}


# ------------------------------------------------------------------ the B200 arm
class Bench:
    """One rank of the B200 arm: device-resident (`value`) and C-ABI (`e2e`) measurements of a workload."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from cvortex_b200 import api
        self.torch, self.dist, self.api, self.args = torch, dist, api, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the all-pairs path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            # a host-side group for waits that must leave the GPUs alone (an NCCL barrier is a kernel that spins
            # on every waiting rank's GPU; without MPS it takes every other time slice from another process's work there)
            self.cpu_group = dist.new_group(backend="gloo")
        api.initialise(require_gpu=True)
        api.use_only(self.local_rank)
        self.be, self.lib = api.backend(), api.library()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        """Rendezvous on the host only: nothing is launched on any GPU while ranks wait here."""
        if self.world > 1:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x, op="max"):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident -------------------------------------------------------------------
    def resident(self, name, steps, warmup, keep_outputs=False, warm_fraction=1.0, sample_clocks=False):
        """`value` of a workload: particles resident in HBM, sharded over the ranks; each step
        all-gathers the raw source rows (N > 1) and runs every op on this rank's targets.  Small
        workloads (< 2e9 pairs) run on rank 0 alone, as the library itself would."""
        torch = self.torch
        from cvortex_b200.sharding import ShardedM2M, target_range
        n, m, ops = WORKLOADS[name]
        single = float(n) * m < SHARD_MIN_PAIRS and self.world > 1
        if single and self.rank != 0:
            return None
        rank, world = (0, 1) if single else (self.rank, self.world)
        P, TP, X = make_inputs(n, m, ops)
        lo, hi = target_range(m, rank, world)
        slo, shi = target_range(n, rank, world)
        m_local = hi - lo
        src_local = torch.from_numpy(np.ascontiguousarray(P[slo:shi])).to(self.dev)
        tgts = {op: torch.from_numpy(np.ascontiguousarray((TP if op in PARTICLE_TARGETS else X)[lo:hi])).to(self.dev) for op, _ in ops}
        outs = {op: torch.empty((m_local, out_cols(op)), device=self.dev) for op, _ in ops}
        sharded = ShardedM2M(self.be, self.local_rank, n)
        st = self.stream
        ev = {op: [] for op, _ in ops}

        def step(record, frac=1.0):
            full = src_local if world == 1 else sharded.gather_sources(src_local)   # NCCL all-gather of the raw rows
            mt = max(1, int(m_local * frac))
            for op, reg in ops:
                if record:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(st)
                self.be.m2m(op, reg, self.local_rank, st.cuda_stream, full, n, tgts[op], mt, outs[op], SIGMA, NU)
                if record:
                    b.record(st)
                    ev[op].append((a, b))
            self.flush.zero_()                                   # evict L2 between steps

        sync = (lambda: torch.cuda.synchronize()) if single else self.barrier
        for _ in range(warmup):
            step(False, warm_fraction)
        sync()
        sampler = ClockSampler(self.local_rank).start() if (sample_clocks and self.rank == 0) else None
        launches0 = self.be.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            step(True)
        e1.record(st)
        sync()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        launches = self.be.kernel_launches() - launches0
        if not single:
            ms = self.max_over_ranks(ms)
            launches = int(self.max_over_ranks(float(launches), "sum"))
        pairs_per_step = float(n) * m * len(ops)
        kern = {}
        for op, reg in ops:
            d = np.array([a.elapsed_time(b) for a, b in ev[op]])
            info = self.be.op_info(op, reg)
            kern[op] = {"reg": reg, "ms": float(d.mean()), "lane_ops": info["lane_ops"], "sfu_ops": info["sfu_ops"],
                        "pairs": float(n) * m_local}
        res = {"value": pairs_per_step * steps / (ms * 1e-3) / 1e9, "ms_per_step": ms / steps, "steps": steps,
               "launches": launches, "kern": kern, "pairs_per_step": pairs_per_step, "n_gpus": 1 if single else self.world,
               "clocks": clocks}
        if keep_outputs:
            res["outs"], res["tgts"], res["src_local"], res["sharded"], res["m_local"] = outs, tgts, src_local, sharded, m_local
        return res

    # ---- through the reference's ABI, host pointer arrays ----------------------------------
    def e2e(self, name, steps):
        from cvortex_b200.abi import PointerRows
        from cvortex_b200.sharding import target_range
        n, m, ops = WORKLOADS[name]
        single = float(n) * m < SHARD_MIN_PAIRS and self.world > 1
        if single and self.rank != 0:
            return None
        rank, world = (0, 1) if single else (self.rank, self.world)
        P, TP, X = make_inputs(n, m, ops)
        lo, hi = target_range(m, rank, world)
        m_local = hi - lo
        # what a C caller holds between calls: the struct arrays, the arrays of pointers into them
        # (built once, as the reference's bench setup does) and preallocated result arrays
        srcs = PointerRows(P, P.shape[1])
        host_t, host_o = {}, {}
        for op, _ in ops:
            rows = np.ascontiguousarray((TP if op in PARTICLE_TARGETS else X)[lo:hi])
            host_t[op] = PointerRows(rows, rows.shape[1]) if op in PARTICLE_TARGETS else rows
            host_o[op] = np.empty((m_local, out_cols(op)), dtype=np.float32)

        def e2e_step():
            res = []
            for op, reg in ops:
                fn = getattr(self.lib, op)
                if op.startswith("F3D"):
                    res.append(fn(srcs, host_t[op], out=host_o[op]))
                elif op.endswith("visc_dvort"):
                    res.append(fn(srcs, host_t[op], reg, SIGMA, NU, out=host_o[op]))
                else:
                    res.append(fn(srcs, host_t[op], reg, SIGMA, out=host_o[op]))
                assert self.be.last_dispatch() == 1
            return res
        e2e_step()
        e2e_step()
        sync = (lambda: self.torch.cuda.synchronize()) if single else self.barrier
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = e2e_step()
        dt = time.perf_counter() - t0
        if not single:
            dt = self.max_over_ranks(dt)
        h2d = sum(P.nbytes + (host_t[op].rows if op in PARTICLE_TARGETS else host_t[op]).nbytes for op, _ in ops)
        d2h = sum(r.nbytes for r in res)
        line = {"value": float(n) * m * len(ops) * steps / dt / 1e9, "unit": "Gpair/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt / steps,
                "api": "cvtx_*_M2M_* C ABI, host arrays of pointers; per rank when N > 1"}
        if single or (self.world == 1 and float(n) * m <= 1e9):
            # small calls: the same loop with result arrays the caller page-locked (the library then has them written
            # directly instead of copying them out of its staging area); the headline figure above stays the pageable one
            for op, _ in ops:
                host_o[op] = self.torch.empty((m_local, out_cols(op)), dtype=self.torch.float32).pin_memory().numpy()
            e2e_step()
            sync()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            line["ms_per_step_page_locked_result"] = 1e3 * (time.perf_counter() - t0) / steps
        return line

    # ---- ONE process, every GPU of the box, through the reference's ABI ------------------------
    def inlib_multi_gpu(self, name, steps):
        """What a C / Julia caller gets from one process: every accelerator switched on with the reference's own
        cvtx_accelerator_enable, host arrays of pointers in, host arrays out.  Inside the library the sources
        cross PCIe once in total (device g uploads rows [g n/G, (g+1) n/G)), the shards are all-gathered over
        NCCL / NVLink, every device runs the pair kernel on its target shard.  Called on rank 0 only, between
        two host-side barriers, while the other ranks idle with nothing queued on their GPUs."""
        from cvortex_b200.abi import PointerRows
        n, m, ops = WORKLOADS[name]
        n_acc = min(self.lib.num_accelerators(), self.world)
        P, TP, X = make_inputs(n, m, ops)
        srcs = PointerRows(P, P.shape[1])
        host_t, host_o = {}, {}
        for op, _ in ops:
            rows = np.ascontiguousarray(TP if op in PARTICLE_TARGETS else X)
            host_t[op] = PointerRows(rows, rows.shape[1]) if op in PARTICLE_TARGETS else rows
            host_o[op] = np.empty((m, out_cols(op)), dtype=np.float32)

        def step():
            for op, reg in ops:
                fn = getattr(self.lib, op)
                if op.startswith("F3D"):
                    fn(srcs, host_t[op], out=host_o[op])
                elif op.endswith("visc_dvort"):
                    fn(srcs, host_t[op], reg, SIGMA, NU, out=host_o[op])
                else:
                    fn(srcs, host_t[op], reg, SIGMA, out=host_o[op])
                assert self.be.last_dispatch() == 1
        try:
            for k in range(n_acc):
                self.lib.accelerator_enable(k)
            step()
            used = self.be.last_devices_used()
            step()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            dt = time.perf_counter() - t0
        finally:
            self.api.use_only(self.local_rank)
        # outside the timed region: the same calls on THIS rank's GPU alone for a stride sample of the targets.  A target's
        # result does not depend on which other targets share the call or on how many devices took part
        # (tests/test_gpu_multidevice.py; skipped on one-GPU boxes, so the driver's scaling run carries the check too).
        idx = np.arange(0, m, max(1, m // 4096))
        same = True
        for op, reg in ops:
            rows = (TP if op in PARTICLE_TARGETS else X)[idx]
            rows = np.ascontiguousarray(rows)
            tg = PointerRows(rows, rows.shape[1]) if op in PARTICLE_TARGETS else rows
            fn = getattr(self.lib, op)
            if op.startswith("F3D"):
                one = fn(srcs, tg)
            elif op.endswith("visc_dvort"):
                one = fn(srcs, tg, reg, SIGMA, NU)
            else:
                one = fn(srcs, tg, reg, SIGMA)
            same = same and bool(np.array_equal(np.asarray(one).reshape(len(idx), -1).view(np.uint32),
                                                host_o[op][idx].reshape(len(idx), -1).view(np.uint32)))
        return {"value": float(n) * m * len(ops) * steps / dt / 1e9, "unit": "Gpair/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
                "devices_used": used, "exchange": self.be.exchange_backend(),
                "bits_equal_to_one_gpu": same, "checked_targets": int(len(idx)),
                "api": "cvtx_*_M2M_* C ABI from ONE process, all accelerators enabled, host arrays of pointers; "
                       "sharded upload + NCCL all-gather of the source rows inside libcvortex.so"}

    # ---- roofline of a measured workload ---------------------------------------------------
    def peaks(self):
        if not hasattr(self, "_peaks"):
            measured = {}
            try:
                measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except (OSError, ValueError):
                pass
            sms = self.be.sm_count(self.local_rank)
            sm_max_mhz = float(measured.get("sm_max_mhz", self.be.clock_khz(self.local_rank) / 1e3))
            self._peaks = {"sms": sms, "sm_max_mhz": sm_max_mhz, "lane": sms * 128 * sm_max_mhz * 1e6,
                           "sfu": sms * 16 * sm_max_mhz * 1e6,
                           "lane_measured": self.be.measure_peak(self.local_rank, "fp32"),
                           "sfu_measured": self.be.measure_peak(self.local_rank, "mufu")}
        return self._peaks

    def kernel_fractions(self, kern):
        pk = self.peaks()
        out = {}
        for k, v in kern.items():
            rate = v["pairs"] / (v["ms"] * 1e-3)
            out[f"{k}/{v['reg']}"] = {"avg_launch_ms": v["ms"], "gpairs_per_s": rate / 1e9,
                                      "lane_ops_per_pair": v["lane_ops"], "sfu_ops_per_pair": v["sfu_ops"],
                                      "frac_fp32": rate * v["lane_ops"] / pk["lane"], "frac_sfu": rate * v["sfu_ops"] / pk["sfu"]}
        return out

    def roofline(self, kern):
        pk = self.peaks()
        dom = max(kern, key=lambda k: kern[k]["ms"])
        kd = kern[dom]
        rate = kd["pairs"] / (kd["ms"] * 1e-3)
        bound = "sfu" if kd["sfu_ops"] * 8 > kd["lane_ops"] else "fp32"
        if bound == "fp32":
            achieved, peak, peak_m, unit = rate * kd["lane_ops"] * 2 / 1e12, pk["lane"] * 2 / 1e12, pk["lane_measured"] * 2 / 1e12, "TFLOP/s"
        else:
            achieved, peak, peak_m, unit = rate * kd["sfu_ops"] / 1e12, pk["sfu"] / 1e12, pk["sfu_measured"] / 1e12, "T MUFU-op/s"
        traffic, traffic_source = None, None
        for tname in ("traffic_r2.json", "traffic_r1.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(f"{dom}/{kd['reg']}")
                if traffic is not None:
                    traffic_source = f"static: dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture of this launch, profiles/{tname} (not measured in this run)"
                    break
        return {
            "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_source,
            "peak_measured": peak_m, "frac_of_measured_peak": achieved / peak_m if peak_m else None,
            "kernel": f"m2m_kernel<{dom}/{kd['reg']}>", "avg_launch_ms": kd["ms"], "pairs_per_launch": kd["pairs"],
            "lane_ops_per_pair": kd["lane_ops"], "sfu_ops_per_pair": kd["sfu_ops"],
            "peak_source": (f"`peak` = nominal FP32 issue peak = {pk['sms']} SMs x 128 lanes x {pk['sm_max_mhz']:.0f} MHz (sm_max_mhz of "
                            "MEASURED_PEAKS.json, which holds HBM and bf16 figures only) x 2 flop, each algorithmic FP32 lane-op "
                            "counted as one FMA slot; `peak_measured` = the library's FFMA2 (resp. MUFU.RSQ) loop timed in this "
                            "run on this GPU (cvtx_b200_measure_peak)"),
            "pipe_peaks_measured": {"fp32_lane_ops_per_s": pk["lane_measured"], "mufu_ops_per_s": pk["sfu_measured"],
                                    "fp32_vs_nominal": pk["lane_measured"] / pk["lane"], "mufu_vs_nominal": pk["sfu_measured"] / pk["sfu"]},
            "all_kernels": self.kernel_fractions(kern),
        }


def parity_sample(torch, dev, ref, outs):
    """The GPU results of a workload (full device-resident outputs of its last step) on the reference leg's own
    target sample, against the reference's outputs (already computed by ref.pass_()) and against the FP64 oracle."""
    idx_t = torch.from_numpy(ref.idx).to(dev)
    parity = {"targets": int(len(ref.idx)), "sample": "all sources x stride-sampled targets, the same sample the CPU leg ran",
              "tolerance": 1e-5, "per_op": {}}
    worst_ref, worst_f64 = 0.0, 0.0
    for op, reg in ref.ops:
        got = outs[op][idx_t].cpu().numpy().reshape(len(ref.idx), -1)
        f64 = ref.f64(op, reg)
        e_ref, e_f64, r_f64 = rel_l2(got, ref.outputs[op]), rel_l2(got, f64), rel_l2(ref.outputs[op], f64)
        parity["per_op"][f"{op}/{reg}"] = {"rel_l2_vs_ref": e_ref, "rel_l2_vs_f64": e_f64, "ref_rel_l2_vs_f64": r_f64,
                                           "finite": bool(np.all(np.isfinite(got)))}
        worst_ref, worst_f64 = max(worst_ref, e_ref), max(worst_f64, e_f64)
    parity["rel_l2_vs_ref"], parity["rel_l2_vs_f64"] = worst_ref, worst_f64
    # within tolerance of the reference, or -- where the FP32 reference is itself further than that from FP64
    # (filaments with short segments) -- at least as close to FP64 as the reference is
    parity["ok"] = bool(all(v["finite"] and (v["rel_l2_vs_ref"] <= 1e-5 or v["rel_l2_vs_f64"] <= 1.1 * v["ref_rel_l2_vs_f64"] + 5e-7)
                            for v in parity["per_op"].values()))
    return parity


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the process's real stdout (see main)."""
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT, text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs / headline extra keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line at the first
    # communicator when the box exports NCCL_DEBUG): everything this process and its libraries print goes to stderr,
    # and only emit() below writes to the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    B = Bench(args)
    torch, be = B.torch, B.be
    n, m, ops = WORKLOADS[args.workload]
    main_res = B.resident(args.workload, args.steps, args.warmup, keep_outputs=True, sample_clocks=True)
    clocks = main_res["clocks"]

    # ---- config 2 "as fused pass" (SURVEY 8d): velocity AT the particles + stretching in one
    # sweep over the sources (thin-ABI op CVTX_B200_P3D_VEL_DVORT); counted as 2 pair-interactions
    # per (source, target) like the two separate ops it replaces
    fused = None
    if [o for o, _ in ops] == ["P3D_M2M_vel", "P3D_M2M_dvort"]:
        reg = ops[0][1]
        m_local, sharded, src_local = main_res["m_local"], main_res["sharded"], main_res["src_local"]
        tg = main_res["tgts"]["P3D_M2M_dvort"]
        out6 = torch.empty((m_local, 6), device=B.dev)
        st = B.stream
        full = sharded.gather_sources(src_local)
        for _ in range(2):
            be.m2m("P3D_M2M_vel_dvort", reg, B.local_rank, st.cuda_stream, full, n, tg, m_local, out6, SIGMA, NU)
        B.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(st)
        for _ in range(args.steps):
            full = sharded.gather_sources(src_local)
            be.m2m("P3D_M2M_vel_dvort", reg, B.local_rank, st.cuda_stream, full, n, tg, m_local, out6, SIGMA, NU)
            B.flush.zero_()
        f1.record(st)
        B.barrier()
        fms = B.max_over_ranks(f0.elapsed_time(f1)) / args.steps
        finfo = be.op_info("P3D_M2M_vel_dvort", reg)
        fused = {"value": 2.0 * n * m / (fms * 1e-3) / 1e9, "unit": "Gpair/s", "ms_per_step": fms,
                 "lane_ops_per_source_target": finfo["lane_ops"],
                 "note": "vel evaluated at the particle positions (not the independent point cloud of the "
                         "separate-op step) fused with dvort; additive thin-ABI op, not part of `value`"}
        del out6

    e2e = None if args.no_e2e else B.e2e(args.workload, args.steps)

    # ---- the same workload from ONE process driving all N GPUs through the library (N > 1)
    inlib = None
    if world > 1 and not args.no_e2e:
        B.barrier()
        B.host_barrier()
        if rank == 0:
            inlib = B.inlib_multi_gpu(args.workload, max(1, min(args.steps, 3)))
        B.host_barrier()                 # the other ranks wait on the host: their GPUs are rank 0's for this measurement

    # ---- cpu_baseline + in-run parity on the same target sample (N = 1 only)
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        ref = ReferenceCpu(n, m, ops, CPU_SAMPLE_TARGETS if n >= 100_000 else m)
        ref.warm()
        ref.pass_()                      # one untimed full pass, as the reference arm does: the first pass runs 5-7 % slow
        secs = ref.pass_()
        cpu = {"value": ref.pairs / secs / 1e9, "unit": "Gpair/s", "cores": ref.cores, "kind": ref.kind,
               "sample": ref.describe(secs)}
        parity = parity_sample(torch, B.dev, ref, main_res["outs"])
        del ref
    main_kern = main_res["kern"]
    main_launches = main_res["launches"]
    main_value, main_ms = main_res["value"], main_res["ms_per_step"]
    for k in ("outs", "tgts", "src_local", "sharded"):
        main_res.pop(k, None)
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and the north-star headline, same method, fewer steps
    extra = {}
    if not args.no_extra:
        plan = [("p3d_vel_winckelmans_1M", min(args.steps, 5), 3, 1.0, True),
                ("p3d_vel_winckelmans_10k", 50, 10, 1.0, True),
                ("f3d_vel+dvort_100k_on_2M", min(args.steps, 3), 3, 1.0, False),
                ("p3d_visc_winckelmans_4M", 1, 3, 1.0 / 16, False),
                ("p2d_vel+visc_gaussian_4M", 1, 3, 1.0 / 16, False)]
        for name, steps, warm, wfrac, with_e2e in plan:
            if name == args.workload:
                continue
            want_parity = world == 1 and not args.no_cpu_baseline
            r = B.resident(name, steps, warm, warm_fraction=wfrac, keep_outputs=want_parity)
            e = B.e2e(name, steps) if (with_e2e and not args.no_e2e) else None
            if r is None:
                continue
            entry = {"value": r["value"], "unit": "Gpair/s", "ms_per_step": r["ms_per_step"], "steps": steps, "n_gpus": r["n_gpus"],
                     "config": workload_config(name, *WORKLOADS[name], world=r["n_gpus"]),
                     "kernels": B.kernel_fractions(r["kern"]) if rank == 0 else None, "e2e": e}
            if want_parity:
                # a smaller sample than the headline's 4096 targets: the 4M-source configs cost the CPU 4 ns per pair
                en, em, eops = WORKLOADS[name]
                xref = ReferenceCpu(en, em, eops, CPU_EXTRA_SAMPLE_TARGETS if en >= 100_000 else em)
                xref.pass_()
                entry["parity"] = parity_sample(torch, B.dev, xref, r["outs"])
                del xref
                for k in ("outs", "tgts", "src_local", "sharded"):
                    r.pop(k, None)
            if wfrac < 1.0:
                entry["warmup_note"] = f"{warm} warm-up passes on 1/{int(round(1 / wfrac))} of this rank's targets (a full pass takes ~10 s)"
            extra[name] = entry
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": "pair-interactions/s", "value": main_value, "unit": "Gpair/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, n, m, ops, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(main_launches),
            "roofline": B.roofline(main_kern), "cpu_baseline": cpu, "parity": parity, "fused_vel_dvort": fused,
            "inlib_multi_gpu": inlib,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
