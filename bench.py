#!/usr/bin/env python
"""Benchmark of the batched multi-rate MPC hot path (BASELINE.json: "MPC solves/sec").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

One *step* = one controller tick of a batch of B = 1024 independent MPC instances per GPU
(BASELINE.json configs[1]: perturbed initial states, reference horizon 17 knots and the
variable-sampling grid): linearise kernel + structured QP kernel + output extraction.

* ``value``  : solves/s with the input packs already resident in HBM (CUDA-event timed on the
               launching stream, L2 flushed between steps, max over ranks).
* ``e2e``    : the same metric through the public C-ABI call sequence with HOST buffers:
               vsmpc_set_state (H2D from pinned memory + K1) -> vsmpc_solve (K2) -> vsmpc_get_output (D2H).
* ``roofline``: FP64 (CUDA-core DFMA) roofline of the QP kernel: algorithmic flops per SURVEY §8(d)
               (F = F_l + n_f F_f + n_s F_s at (nx,nu,N) = (30,12,17)) over the kernel's own CUDA-event
               time, against the DFMA peak measured on this device in the same run.
* ``cpu_baseline``: the oracle port timed on this box's host cores on a bounded sample.

Multi-GPU: instances are independent — each rank owns its own batch (weak scaling), no collective in
the data path; results are only gathered (timing max over ranks).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200"
METRIC = "MPC solves/sec (batched)"
B_PER_GPU = 1024


def pkg(sub=""):
    return importlib.import_module(PKG + (("." + sub) if sub else ""))


# ---- algorithmic flop model, SURVEY.md §8(d) / BASELINE.md §5 --------------------------------------
def flops_per_solve(nx, nu, N, n_f, n_s):
    F_f = N * (7.0 / 3.0 * nx ** 3 + 4 * nx ** 2 * nu + 2 * nx * nu ** 2 + nu ** 3 / 3.0)
    F_s = N * (8 * nx ** 2 + 8 * nx * nu + 2 * nu ** 2)
    F_l = 4000 + 2 * N * (nx ** 2 + nx * nu + nx)
    return F_l + n_f * F_f + n_s * F_s, dict(F_f=F_f, F_s=F_s, F_l=F_l)


_NVML_CHILD = r"""
import os, sys, time
import pynvml as N
N.nvmlInit()
ppid = os.getppid()
ident, path = sys.argv[1], sys.argv[2]
try:
    h = N.nvmlDeviceGetHandleByUUID(ident if ident.startswith("GPU-") else "GPU-" + ident) if "-" in ident else N.nvmlDeviceGetHandleByIndex(int(ident))
except Exception:
    h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[3]))
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
with open(path, "w") as f:
    while True:
        f.write("%.6f,%d,%d,%d\n" % (time.time(), N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), mx, int(reasons(h))))
        f.flush()
        time.sleep(0.004)
        if os.getppid() != ppid:      # the bench is gone: do not linger
            break
"""


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (B200_PROFILING.md's clocks line).  The timed region of
    this bench is tens of milliseconds, shorter than one `nvidia-smi -lms` period plus its start-up, so the sampler is a child
    process reading NVML every 4 ms into a file (a Python thread in this process would take the GIL away from the host loop
    being timed); the parent marks the timed windows with wall-clock times and summarises the samples that fall inside them.
    Falls back to `nvidia-smi -lms 20` if NVML's Python module is missing."""
    # NVML clocks event reasons (nvml.h): sw power cap 0x4, hw slowdown 0x8, sw thermal 0x20, hw thermal 0x40
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index: int, uuid: str = ""):
        import tempfile
        self.gpu, self.uuid = gpu_index, uuid
        self.samples, self.windows = [], []
        self.proc, self.mode = None, None
        self.path = tempfile.mktemp(prefix="vsmpc_clocks_", suffix=".csv")

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_CHILD, self.uuid or str(self.gpu), self.path, str(self.gpu)],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            self.mode = "nvml"
        except Exception:
            self.proc = None

    def ready(self):
        """True once the child has written its first sample (or has died: then the nvidia-smi fallback is started)."""
        try:
            if os.path.getsize(self.path) > 0:
                return True
        except OSError:
            pass
        if self.proc is not None and self.proc.poll() is not None and self.mode == "nvml":
            q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown")
            try:
                self.out = open(self.path, "w")
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.uuid and 'GPU-' + self.uuid or self.gpu}", f"--query-gpu={q}",
                                              "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.out, stderr=subprocess.DEVNULL)
                self.mode = "smi"
            except Exception:
                self.proc, self.mode = None, "none"
                return True
        return self.proc is None

    def begin(self):
        self.windows.append([time.time(), None])

    def end(self):
        if self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.time()

    def stop(self):
        if self.proc:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                pass
        try:
            with open(self.path) as f:
                self.samples = [[x.strip() for x in line.split(",")] for line in f if line.strip()]
            os.remove(self.path)
        except Exception:
            pass

    def summary(self):
        sm, mx, reasons, sm_all = [], [], set(), []
        for s in self.samples:
            try:
                if self.mode == "nvml":
                    t, c, m, r = float(s[0]), float(s[1]), float(s[2]), int(s[3])
                    # windows padded by 20 ms: the GPU is under the same load right before (warm-up) and after them
                    inside = any(a - 0.02 <= t <= (b if b is not None else t) + 0.02 for a, b in self.windows) or not self.windows
                    rs = [n for n, bit in self.BITS if r & bit]
                else:       # nvidia-smi lines carry no epoch time: every line counts (the loop started before the windows)
                    c, m = float(s[1]), float(s[2])
                    inside = True
                    rs = [n for (n, _), v in zip(self.BITS, s[3:7]) if v.lower().startswith("active")]
                sm_all.append(c)
                if inside:
                    sm.append(c)
                    mx.append(m)
                    reasons.update(rs)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.mode}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_min": float(min(sm)), "source": "NVML every 4 ms, samples inside the timed windows (+- 20 ms)" if self.mode == "nvml"
                else "nvidia-smi -lms 20", "samples_total": len(sm_all)}


def load_traj():
    return pkg("config").load_trajectories_npz(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))


def make_workload(B, seed, n_sets):
    """n_sets distinct perturbed packs (SURVEY §8(d) Config 2) + the nominal configure pack."""
    syn, pack = pkg("synthetic"), pkg("pack")
    nom = syn.make_states(B, perturbed=False)
    packs = [pack.build_pack(syn.make_states(B, seed=seed + 7919 * j, perturbed=True)) for j in range(n_sets)]
    jp = np.ascontiguousarray(nom["joint_pos"][:, pack.DEFAULT_JOINT_SELECTOR].T)
    return pack.build_pack(nom), jp, packs


# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bat, L = pkg("batched"), pkg("_lib")
    lib = L.load()
    B = args.batch
    K, Wm = args.steps, args.warmup
    n_sets = 4
    nom_pack, jp, packs = make_workload(B, 20251002 + rank, n_sets)
    params = None
    if args.horizon:     # BASELINE configs[3]: long-horizon variant (the default run is configs[1], reference horizon)
        hN, hNs, hNc = [int(x) for x in args.horizon.split(",")]
        params = dict(nIter=hN, nIterSmall=hNs, controlHorizon=hNc)
        args.no_latency, args.rollout_ticks, args.no_cpu = True, 0, True
    mpc = bat.BatchedVSMPC(B, params, load_traj(), device=local_rank, solver=args.solver)
    stream = torch.cuda.Stream(device=dev)   # an explicit stream: events and kernels share it
    torch.cuda.set_stream(stream)
    mpc.set_stream(stream.cuda_stream)
    # stagger the 20-tick phase so that every step has the steady-state mix of pinned / released ticks
    phase0 = (np.arange(B) % 20).astype(np.int32)
    mpc.configure_pack(nom_pack, jp, phase0)
    d_packs = [torch.from_numpy(p).to(dev) for p in packs]
    h_packs = [torch.from_numpy(p).pin_memory() for p in packs]
    h_out = torch.empty((B, L.OUT_DOUBLES), dtype=torch.float64).pin_memory()
    h_status = torch.empty((B,), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def step_dev(j):
        mpc.update_device_ptr(d_packs[j % n_sets].data_ptr())
        mpc.solve_async()

    def step_e2e(j):
        mpc.update_ptr(h_packs[j % n_sets].data_ptr())
        mpc.solve_async()
        mpc.get_output_into(h_out.data_ptr(), h_status.data_ptr())  # blocks until the D2H copy landed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # host buffers of the end-to-end leg, built BEFORE any timed leg: generating them takes seconds of host time, and a GPU
    # left idle that long drops its clocks — the e2e leg would then be timed while they ramp up again
    pack_bytes = packs[0].nbytes
    n_e2e = max(n_sets, int(np.ceil(140e6 / pack_bytes)))
    syn, packm = pkg("synthetic"), pkg("pack")
    h_e2e = list(h_packs)
    for j in range(n_sets, n_e2e):
        h_e2e.append(torch.from_numpy(packm.build_pack(
            syn.make_states(B, seed=20251002 + rank + 7919 * j, perturbed=True))).pin_memory())
    h_out2 = [torch.empty((B, L.OUT_DOUBLES), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_status2 = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    # the first DMA out of a freshly pinned buffer is slower than the following ones (its pages are mapped for the device on
    # first use: +17 us per 2.9 MB pack, profiles/r02z_e2e.txt): touch every host buffer once, as part of allocating it
    scratch = torch.empty_like(d_packs[0])
    for t in h_e2e:
        scratch.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    del scratch

    # ---------------- device-resident leg -------------------------------------------------------------
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = ""
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()
    for j in range(Wm):
        step_dev(j)
    # the warm-up goes on (GPU under load, clocks up) until the clock sampler has delivered its first sample
    t_ready = time.perf_counter()
    j = Wm
    while not sampler.ready() and time.perf_counter() - t_ready < 5.0:
        step_dev(j)
        j += 1
        if j % 16 == 0:
            torch.cuda.synchronize()
    barrier()
    sampler.begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    nf_sum = ns_sum = 0.0
    t_wall0 = time.perf_counter()
    for j in range(K):
        flush.zero_()                      # L2 flush between timed iterations (outside the events)
        e0, e1, e2 = ev[j]
        e0.record(stream)
        mpc.update_device_ptr(d_packs[j % n_sets].data_ptr())
        e1.record(stream)
        mpc.solve_async()
        e2.record(stream)
    barrier()
    sampler.end()
    t_wall = time.perf_counter() - t_wall0
    k1_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in ev)
    k2_ms = sum(e1.elapsed_time(e2) for _, e1, e2 in ev)
    total_ms = sum(e0.elapsed_time(e2) for e0, _, e2 in ev)
    nf, ns = mpc.get_counts()
    _, status = mpc.get_output()
    solved_frac = float((status == 0).mean())
    # ---------------- end-to-end leg (host buffers through the C-ABI) -----------------------------------
    # every step: vsmpc_set_state (H2D of THAT step's pack from pinned host memory, staged on the library's copy
    # stream) -> vsmpc_solve_async -> vsmpc_get_output_async (D2H of the step's 54-double rows + status) and the host
    # waits for the result of the step before: two steps in flight, the copy of step j+1 overlaps the QP kernel of
    # step j.  The host packs cycle through more bytes than the 126 MB L2 (no flush kernel in this loop).
    def e2e_loop(n):
        prev = None
        for j in range(n):
            mpc.update_ptr(h_e2e[j % n_e2e].data_ptr())
            mpc.solve_async()
            t = mpc.get_output_async(h_out2[j & 1].data_ptr(), h_status2[j & 1].data_ptr())
            if prev is not None:
                mpc.wait_output(prev)
            prev = t
        mpc.wait_output(prev)

    e2e_loop(Wm)
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    e2e_loop(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.end()
    barrier()
    # the same sequence with a blocking read-back every step (no overlap), for reference
    for j in range(Wm):
        step_e2e(j)
    barrier()
    e2e_blocking_s = 0.0
    for j in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_e2e(j)
        e2e_blocking_s += time.perf_counter() - t0
    barrier()
    sampler.stop()
    # ---------------- single-instance latency (BASELINE metric: "single-solve p50 latency") -----------------
    latency = None
    if rank == 0 and not args.no_latency:
        latency = single_solve_latency(bat, L, local_rank, stream, n_ticks=400)
    # ---------------- closed loop on the device (configs[2] style): plant + K1 + K2 per tick, no host round trip ----
    closed = None
    if rank == 0 and args.rollout_ticks > 0:
        closed = closed_loop_leg(bat, B, local_rank, stream, args.rollout_ticks, args.solver)
    # ---------------- the other BASELINE configurations, every rank (skipped by --no-extras / --horizon) ------------
    gather = long_h = monte = sweep = e2e_kin = cpp_host = None
    if not args.no_extras and not params:
        if rank == 0:
            cpp_host = cpp_host_leg(B, nom_pack, jp, packs, max(K, 50), Wm, local_rank)
        e2e_kin = e2e_kinematics_leg(bat, L, B, world, rank, local_rank, stream, dev, K, Wm)
        gather = gather_leg(mpc, B, world, dev, stream, h_out, h_status)
        long_h = long_horizon_leg(bat, lib, d_packs, nom_pack, jp, phase0, B, local_rank, stream, flush, world, dev, args.solver)
        monte = monte_carlo_leg(bat, args.mc_instances, args.mc_ticks, rank, world, local_rank, stream, dev)
        sweep = monte_carlo_leg(bat, args.sweep_instances, args.mc_ticks, rank, world, local_rank, stream, dev, joint_boxes=True)
    # ---------------- reduce over ranks: max time --------------------------------------------------------
    t = torch.tensor([total_ms, e2e_s * 1e3, k1_ms, k2_ms, e2e_blocking_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, k1_ms, k2_ms, e2e_blk_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * K / (total_ms * 1e-3)
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    # ---------------- roofline of the QP kernel -----------------------------------------------------------
    import ctypes
    tf = ctypes.c_double(0.0)
    lib.vsmpc_microbench_fp64(local_rank, 0, ctypes.byref(tf))
    tf_dmma = ctypes.c_double(0.0)
    lib.vsmpc_microbench_fp64(local_rank, 1, ctypes.byref(tf_dmma))
    nf_mean, ns_mean = float(nf.mean()), float(ns.mean())
    F, parts = flops_per_solve(30, 12, params["nIter"] if params else 17, nf_mean, ns_mean)
    k2_s_per_launch = k2_ms * 1e-3 / K
    achieved = F * B / k2_s_per_launch / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not params:     # the committed capture is of the reference-horizon kernel
        try:
            traffic = json.load(open(tpath)).get("qp_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                "frac": achieved / tf.value if tf.value > 0 else None, "traffic": traffic,
                "kernel": ("qp_condensed_wide_kernel" if params and args.solver == 0 else
                           {0: "qp_condensed_kernel", 1: "qp_generic_kernel", 2: "qp_structured_kernel"}[args.solver]),
                "peak_source": "DFMA microbenchmark on this device in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "dmma_peak_tflops": tf_dmma.value,
                "flops_per_solve": F, "n_factor_mean": nf_mean, "n_solve_mean": ns_mean,
                "kernel_ms_per_launch": k2_s_per_launch * 1e3, "linearise_ms_per_launch": k1_ms / K,
                "hbm_algorithmic_bytes_per_solve": 359 * 8 + 54 * 8,
                "hbm_gbs_at_value": value * (359 * 8 + 54 * 8) / 1e9}
    cpu = None if args.no_cpu else cpu_baseline(
        sample_solves=max(args.cpu_sample, min(1024, 4 * (os.cpu_count() or 1))), ticks=5)
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("configs[3]: long-horizon variant, %d knots (%d fine), control horizon %d, batch per GPU"
                                % (hN, hNs, hNc)) if params else
                               "configs[1]: batch of 1024 independent MPC solves per GPU, reference horizon "
                               "(17 knots, 7 fine + 10 coarse), perturbed states (SURVEY §8d Config 2)",
                   "instances_per_gpu": B, "n_var": mpc.n_var, "n_con": mpc.n_con,
                   "l2": "value leg: flushed (256 MiB memset) between timed steps; e2e leg: host inputs cycle through > 126 MB",
                   "pack_sets": n_sets,
                   "solver": {0: "condensed", 1: "generic-dense", 2: "structured"}[args.solver],
                   "phase": "20-tick phase staggered across instances"},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(packs[0].nbytes),
                "d2h_bytes_per_step": int(h_out.numel() * 8 + h_status.numel() * 4),
                "ms_per_step": e2e_ms / K,
                "timing": "host perf_counter around K pipelined steps (set_state + solve_async + get_output_async, two "
                          "steps in flight), synchronized on both sides",
                "host_pack_sets": n_e2e, "host_pack_bytes_cycled": int(n_e2e * pack_bytes),
                "solved_fraction_last_step": float((h_status2[(K - 1) & 1].numpy() == 0).mean()),
                "blocking_value": world * B * K / (e2e_blk_ms * 1e-3),
                "blocking_note": "same calls with a blocking read-back every step (no copy/compute overlap)"},
        # kernels of this repository launched inside the device-timed region: linearise + QP (+ the fallback kernel behind the
        # condensed kernels: it runs every tick and finds its list empty on this workload)
        "gpu_launches": (3 if args.solver == 0 else 2) * K,
        "gpu_launches_per_step": ["linearise_kernel", "qp_condensed_kernel" if not params else "qp_condensed_wide_kernel",
                                  "qp_fallback_kernel"] if args.solver == 0 else 2,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
        "single_solve_latency": latency, "closed_loop": closed,
        "e2e_kinematics": e2e_kin, "e2e_cpp_host": cpp_host, "monte_carlo": monte, "param_sweep": sweep, "long_horizon": finish_long_horizon(long_h, tf.value), "gather": gather,
        "solved_fraction": solved_frac, "wall_ms_timed_loop": t_wall * 1e3,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _max_over_ranks(values, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_leg(mpc, B, world, dev, stream, h_out, h_status, reps=20):
    """Collecting the [B x 54] output rows (the only communication of the path): (a) NCCL all-gather
    straight from the device-resident rows of every rank (vsmpc_get_output_device, no host round trip), (b) every rank
    reading its own rows back into pinned host memory.  CUDA events / host clock, max over ranks."""
    import torch
    sh = pkg("sharding")
    res = {"rows_per_rank": int(B), "row_bytes": 54 * 8 + 4, "reps": reps}
    if world > 1:
        for _ in range(3):
            sh.gather_output_device(mpc, world * B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            rows, st = sh.gather_output_device(mpc, world * B)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = _max_over_ranks([e0.elapsed_time(e1) / reps], world, dev)[0]
        res["nccl_all_gather_ms"] = ms
        res["nccl_all_gather_GBps_per_rank_in"] = (world - 1) * B * (54 * 8 + 4) / (ms * 1e-3) / 1e9
        res["gathered_rows"] = int(rows.shape[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        mpc.get_output_into(h_out.data_ptr(), h_status.data_ptr())
    ms = (time.perf_counter() - t0) * 1e3 / reps
    res["pinned_d2h_ms"] = _max_over_ranks([ms], world, dev)[0]
    res["what"] = ("nccl_all_gather_ms: all_gather_into_tensor of rows + status from the library's device buffers on "
                   "every rank (world > 1 only); pinned_d2h_ms: vsmpc_get_output into pinned host memory per rank")
    return res


def e2e_kinematics_leg(bat, L, B, world, rank, device, stream, dev, K, Wm):
    """End to end with ROBOT STATES as the host input: the kinematics kernel builds the pack on the device
    (Robot::setState, UT/src/Robot.cpp:212-332), so a step uploads 92 doubles per instance instead of 359.  Same pipelined
    call sequence as the e2e leg: vsmpc_set_state_kinematics -> vsmpc_solve_async -> vsmpc_get_output_async."""
    import torch
    kin, syn = pkg("kinematics"), pkg("synthetic")
    model = kin.synthetic_humanoid()
    nd = model["n_dof"]
    q0 = syn.SyntheticRobot().joint_pos0[:nd]
    hover = float(np.sum(model["mass"])) * 9.81 / 4.0

    def states(seed, perturbed):
        g = np.random.default_rng(seed)
        z = (lambda *s: g.normal(0, 1.0, (B,) + s)) if perturbed else (lambda *s: np.zeros((B,) + s))
        return dict(wRb=syn.rpy_to_R(0.05 * z(3)), base_pos=np.array([0.0, 0.0, 1.0]) + 0.05 * z(3), base_lin_vel=0.05 * z(3),
                    omega_world=0.1 * z(3), q=q0[None] + 0.02 * z(nd), qd=0.05 * z(nd), thrust=hover + 8.0 * z(4),
                    thrust_dot_est=20.0 * z(4), thrust_des=hover + 8.0 * z(4), thrust_dot_des=10.0 * z(4),
                    throttle_prev=60.0 + 10.0 * z(4), q_cmd=q0[None] + 0.02 * z(nd))

    mpc = bat.BatchedVSMPC(B, None, load_traj(), device=device)
    mpc.set_stream(stream.cuda_stream)
    fe = kin.KinematicsFrontEnd(mpc, model)
    fe.configure(kin.build_kin_state(model, states(1, False)), (np.arange(B) % 20).astype(np.int32))
    n_sets = 8
    h_ks = [torch.from_numpy(kin.build_kin_state(model, states(20251002 + rank + 31 * j, True))).pin_memory() for j in range(n_sets)]
    scratch = torch.empty_like(h_ks[0], device=dev)
    for t in h_ks:
        scratch.copy_(t, non_blocking=True)
    h_out = [torch.empty((B, L.OUT_DOUBLES), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_st = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    views = [t.numpy() for t in h_ks]

    def loop(n):
        prev = None
        for j in range(n):
            fe.update(views[j % n_sets])
            mpc.solve_async()
            t = mpc.get_output_async(h_out[j & 1].data_ptr(), h_st[j & 1].data_ptr())
            if prev is not None:
                mpc.wait_output(prev)
            prev = t
        mpc.wait_output(prev)

    loop(Wm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(K)
    torch.cuda.synchronize()
    ms = _max_over_ranks([(time.perf_counter() - t0) * 1e3], world, dev)[0]
    solved = float((h_st[(K - 1) & 1].numpy() == 0).mean())
    mpc.close()
    return {"value": world * B * K / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms / K,
            "h2d_bytes_per_step": int(h_ks[0].numel() * 8), "d2h_bytes_per_step": int(B * (L.OUT_DOUBLES * 8 + 4)),
            "gpu_launches_per_step": 3, "solved_fraction_last_step": solved,
            "what": "robot states (base pose and twist, 23 joint positions and velocities, 28 QPInput scalars: 92 doubles per "
                    "instance) from pinned host memory; kinematics kernel + linearise + QP on the device; synthetic 24-link tree"}


def long_horizon_leg(bat, lib, d_packs, nom_pack, jp, phase0, B, device, stream, flush, world, dev, solver, steps=10):
    """BASELINE configs[3]: 2x / 3x / 4x the reference knot count (coarse tail), B instances per GPU on every rank (weak
    scaling), same packs as the main leg (the pack does not depend on the horizon), CUDA events, L2 flushed."""
    import torch
    out = []
    for hN, hNs, hNc in ((34, 14, 24), (51, 14, 36), (68, 14, 48)):
        mpc = bat.BatchedVSMPC(B, dict(nIter=hN, nIterSmall=hNs, controlHorizon=hNc), load_traj(), device=device, solver=solver)
        mpc.set_stream(stream.cuda_stream)
        mpc.configure_pack(nom_pack, jp, phase0)
        for j in range(3):
            mpc.update_device_ptr(d_packs[j % len(d_packs)].data_ptr())
            mpc.solve_async()
        torch.cuda.synchronize()
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
        for j in range(steps):
            flush.zero_()
            e0, e1, e2 = ev[j]
            e0.record(stream)
            mpc.update_device_ptr(d_packs[j % len(d_packs)].data_ptr())
            e1.record(stream)
            mpc.solve_async()
            e2.record(stream)
        torch.cuda.synchronize()
        tot = sum(a.elapsed_time(c) for a, _, c in ev)
        k2 = sum(b.elapsed_time(c) for _, b, c in ev)
        nf, ns = mpc.get_counts()
        _, status = mpc.get_output()
        tot, k2 = _max_over_ranks([tot, k2], world, dev)
        F, _ = flops_per_solve(30, 12, hN, float(nf.mean()), float(ns.mean()))
        out.append({"knots": hN, "fine_knots": hNs, "control_horizon": hNc, "n_var": mpc.n_var, "instances_per_gpu": int(B),
                    "value": world * B * steps / (tot * 1e-3), "unit": "solves/s", "ms_per_step": tot / steps,
                    "kernel_ms_per_launch": k2 / steps, "flops_per_solve": F, "steps": steps,
                    "solved_fraction": float((status == 0).mean())})
        mpc.close()
    return out


def finish_long_horizon(rows, peak_tflops):
    if not rows:
        return None
    for r in rows:
        ach = r["flops_per_solve"] * r["instances_per_gpu"] / (r["kernel_ms_per_launch"] * 1e-3) / 1e12
        r["roofline_frac"] = ach / peak_tflops if peak_tflops > 0 else None
    return {"scaling": "weak", "variants": rows,
            "what": "configs[3]: qp_condensed_wide_kernel, same synthetic packs as the main leg, L2 flushed between steps"}


def cpp_host_leg(B, nom_state_pack, jp_rows, packs, K, Wm, device):
    """north_star (a), the C++ host layer end to end: examples/cpp_host_bench.cpp packs the per-instance records of every tick
    (array of structures) into the page-locked SoA buffer (vsmpc::PackBatch::setMany), uploads, solves and reads the rows back,
    two ticks in flight — the same C-ABI calls as the e2e leg, with the AoS -> SoA pack INSIDE the timed region.  Runs the
    compiled example as a child process on this rank's device; rank 0 only."""
    import tempfile
    exe = os.path.join(ROOT, "examples", "bin", "cpp_host_bench")
    if not os.path.exists(exe):
        return {"unavailable": "examples/bin/cpp_host_bench not built"}
    d = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    pk = pkg("pack")
    # joint positions of every instance (all joints): only the controlled ones matter, take them from the configure rows
    nj = 23
    jpos = np.zeros((B, nj))
    jpos[:, list(pk.DEFAULT_JOINT_SELECTOR)] = jp_rows.T
    blob = np.concatenate([
        [d["alphaGravity"].size, d["positionCoM"].shape[1], len(packs), nj, B],
        d["alphaGravity"].ravel(), d["positionCoM"].T.ravel(), d["velocityCoM"].T.ravel(), d["RPY"].T.ravel(),
        d["RPYDot"].T.ravel(), np.array(list(pk.DEFAULT_JOINT_SELECTOR), dtype=np.float64), jpos.ravel(),
        nom_state_pack.T.ravel()] + [p.T.ravel() for p in packs]).astype(np.float64)          # instance-major records
    path = tempfile.mktemp(prefix="vsmpc_cpp_host_", suffix=".bin")
    blob.tofile(path)
    threads = max(1, min(8, (os.cpu_count() or 2) // 2))
    try:
        res = subprocess.run([exe, path, str(K), str(max(Wm, 3)), str(threads), str(device)], capture_output=True, text=True, timeout=120)
        if res.returncode != 0:
            return {"unavailable": "cpp_host_bench rc %d: %s" % (res.returncode, res.stderr.strip()[-200:])}
        out = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001 - a bench leg must not take the line down
        return {"unavailable": "cpp_host_bench: %s" % e}
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    out["what"] = ("C++ host (include/vsmpc_adapter.hpp): per tick PackBatch::setMany of B per-instance records into page-locked SoA "
                   "+ vsmpc_set_state + vsmpc_solve_async + vsmpc_get_output_async, two ticks in flight; pack inside the timed region")
    return out


def monte_carlo_leg(bat, n_total, ticks, rank, world, device, stream, dev, joint_boxes=False):
    """BASELINE configs[2] x configs[4]: a Monte Carlo closed-loop sweep of n_total instances IN TOTAL (contiguous
    ranges over the ranks: strong scaling) over initial states, constant thrust disturbances and per-instance model
    parameters (jet coefficients / normalisation, mass, inertia, throttle limits), `ticks` controller ticks each,
    entirely on the device (surrogate plant); CUDA events, max over ranks.
    joint_boxes (BASELINE configs[4] "per-instance constraint sets", SURVEY §8d Config 5): n_total instances PER RANK (weak
    scaling: 16 384 on eight GPUs = 2048 each), the same parameter sweep with the optional joint-limit rows on and every instance
    its own box around the commanded posture (two thirds 0.1-0.5 rad to either side, a third open)."""
    import torch
    ro, syn, cfg, sh = pkg("rollout"), pkg("synthetic"), pkg("config"), pkg("sharding")
    lo, hi = (rank * n_total, (rank + 1) * n_total) if joint_boxes else sh.shard_range(n_total, rank, world)
    Bl = hi - lo
    rb = syn.SyntheticRobot()
    g = np.random.default_rng(20251002 + 7 * rank)
    ms_, isc = g.uniform(0.9, 1.1, Bl), g.uniform(0.8, 1.2, Bl)
    st = syn.make_states(Bl, seed=20251002 + 7 * rank, perturbed=True, near_bound_fraction=0.0, mass_scale=ms_, inertia_scale=isc)
    st["thrust"] = (rb.mass * ms_ * 9.81 / 4.0)[:, None] + g.normal(0, 8.0, (Bl, 4))
    st["thrust_des"] = st["thrust"].copy()
    st["thrust_dot_est"] = g.normal(0, 5.0, (Bl, 4))
    st["thrust_dot_des"] = np.zeros((Bl, 4))
    st["throttle_prev"] = np.full((Bl, 4), 76.0) + g.normal(0, 3.0, (Bl, 4))
    st["momentum_body"] *= 0.2
    st["q_cmd"] = np.tile(rb.joint_pos0, (Bl, 1))
    coeff = np.tile(np.asarray(cfg.JET_COEFF), (Bl, 1))
    norm = np.tile(np.asarray(cfg.JET_NORM), (Bl, 1))
    coeff[:, 1] *= g.uniform(0.9, 1.1, Bl)
    coeff[:, 2] *= g.uniform(0.9, 1.1, Bl)
    norm[:, 0] *= g.uniform(0.9, 1.1, Bl)
    norm[:, 1] *= g.uniform(0.9, 1.1, Bl)
    trj = dict(load_traj())
    trj["alphaGravity"] = np.ones_like(trj["alphaGravity"])      # in flight: the surrogate has no ground contact
    params = None
    if joint_boxes:
        pk = pkg("pack")
        params = dict(jointPos_min=[-180.0] * 8, jointPos_max=[180.0] * 8)        # replaced per instance below
    mpc = bat.BatchedVSMPC(Bl, params, trj, device=device)
    mpc.set_stream(stream.cuda_stream)
    mpc.set_instance_params(coeff, norm, g.uniform(0.0, 20.0, Bl), g.uniform(80.0, 100.0, Bl))
    if joint_boxes:
        qc = st["q_cmd"][:, pk.DEFAULT_JOINT_SELECTOR]
        width = np.where(np.arange(Bl)[:, None] % 3 == 0, 3.0, g.uniform(0.1, 0.5, (Bl, 8)))
        mpc.set_joint_limits(qc - width, qc + width * g.uniform(0.5, 1.5, (Bl, 8)))
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, mass_scale=ms_, inertia_scale=isc, thrust_disturbance=g.normal(0, 10.0, (Bl, 4)),
              phase0=(np.arange(Bl) % 20).astype(np.int32))
    loop.run(3)                                     # warm-up + graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    loop.run(ticks)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    _, status = mpc.get_output()
    nf, _ = mpc.get_counts()
    ps = loop.plant_state()
    fin = np.isfinite(ps).all(axis=0)
    drift = np.linalg.norm(ps[0:3].T - st["p_com"], axis=1)
    mpc.close()
    import torch.distributed as dist
    cnt = torch.tensor([float((status == 0).sum()), float(fin.sum()), float(Bl), float(nf.sum()), float((nf >= 2).sum())],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max = _max_over_ranks([ms], world, dev)[0]
    solved, finite, total, nf_sum, nf_multi = [float(x) for x in cnt.tolist()]
    out = {"value": total * ticks / (ms_max * 1e-3), "unit": "closed-loop solves/s", "scaling": "weak" if joint_boxes else "strong",
           "instances_total": int(total), "instances_per_rank": int(Bl), "ticks": int(ticks), "ms_total": ms_max,
           "ms_per_tick": ms_max / ticks, "kernels_per_tick": 3, "cuda_graph": True,
           "solved_fraction_last_tick": solved / total, "finite_plant_states": finite / total,
           "rank0_median_com_drift_m": float(np.median(drift[fin])),
           "what": "per-instance initial state, thrust disturbance N(0, 10 N), mass x U(0.9, 1.1), inertia x U(0.8, 1.2), jet "
                   "c1, c2, mu_T, sigma_T x U(0.9, 1.1), throttleMin U(0, 20), throttleMax U(80, 100); surrogate plant "
                   "(5 x 1 ms) + linearise + QP per tick, device-resident, staggered 20-tick phases"}
    if joint_boxes:
        out["what"] += ("; joint-limit rows on, per-instance boxes around the commanded posture (two thirds 0.1-0.5 rad to either side, a "
                        "third open), carried by the QP kernel's own working set")
        out["factorisations_per_solve_last_tick"] = nf_sum / total
        out["instances_refactorised_last_tick"] = nf_multi / total
    return out


def closed_loop_leg(bat, B, device, stream, n_ticks, solver):
    """B closed loops (surrogate plant) advanced n_ticks controller ticks on the device; CUDA-graph
    replay of the three kernels of a tick; timed with CUDA events on the launching stream."""
    import torch
    syn, ro = pkg("synthetic"), pkg("rollout")
    rb = syn.SyntheticRobot()
    g = np.random.default_rng(20251002)
    st = syn.make_states(B, seed=20251002, perturbed=True, near_bound_fraction=0.0)
    st["thrust"] = np.full((B, 4), rb.mass * 9.81 / 4.0) + g.normal(0, 8.0, (B, 4))
    st["thrust_des"] = st["thrust"].copy()
    st["throttle_prev"] = np.full((B, 4), 76.0) + g.normal(0, 3.0, (B, 4))
    st["momentum_body"] *= 0.2
    st["q_cmd"] = np.tile(rb.joint_pos0, (B, 1))
    trj = dict(load_traj())
    trj["alphaGravity"] = np.ones_like(trj["alphaGravity"])     # in flight: the surrogate has no ground contact
    mpc = bat.BatchedVSMPC(B, None, trj, device=device, solver=solver)
    mpc.set_stream(stream.cuda_stream)
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, thrust_disturbance=g.normal(0, 10.0, (B, 4)), phase0=(np.arange(B) % 20).astype(np.int32))
    loop.run(5)                                     # warm-up + graph capture
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    loop.run(n_ticks)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ps = loop.plant_state()
    _, status = mpc.get_output()
    mpc.close()
    return {"value": B * n_ticks / (ms * 1e-3), "unit": "solves/s", "ticks": int(n_ticks), "instances": int(B),
            "ms_per_tick": ms / n_ticks, "kernels_per_tick": 3, "cuda_graph": True,
            "solved_fraction_last_tick": float((status == 0).mean()),
            "max_abs_com_drift_m": float(np.abs(ps[0:3].T - st["p_com"]).max()),
            "what": "in-flight closed loops (alphaGravity = 1, perturbed states, constant thrust disturbances): surrogate "
                    "plant (5 x 1 ms) + linearise + QP per tick, all device-resident"}


def single_solve_latency(bat, L, device, stream, n_ticks=400):
    """One MPC instance driven tick by tick through the C-ABI with host buffers (the reference controller's
    use: update + solveMPC + getters, src/variable_sampling_mpc.py:110-127): per-tick wall time."""
    import torch
    nom_pack, jp, packs = make_workload(1, 4242, 8)
    mpc = bat.BatchedVSMPC(1, None, load_traj(), device=device)
    mpc.set_stream(stream.cuda_stream)
    mpc.configure_pack(nom_pack, jp, None)
    h_packs = [torch.from_numpy(p).pin_memory() for p in packs]
    h_out = torch.empty((1, L.OUT_DOUBLES), dtype=torch.float64).pin_memory()
    h_status = torch.empty((1,), dtype=torch.int32).pin_memory()
    ts = []
    for j in range(n_ticks + 20):
        t0 = time.perf_counter()
        mpc.update_ptr(h_packs[j % len(h_packs)].data_ptr())
        mpc.solve_async()
        mpc.get_output_into(h_out.data_ptr(), h_status.data_ptr())
        ts.append(time.perf_counter() - t0)
    mpc.close()
    ts = np.array(ts[20:]) * 1e3
    return {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)),
            "mean_ms": float(ts.mean()), "ticks": int(n_ticks), "controller_period_ms": 5.0,
            "reference_published_ms": 2.18,
            "what": "vsmpc_set_state (H2D) + vsmpc_solve + vsmpc_get_output (D2H), B = 1, host wall clock"}


# =====================================================================================================
def cpu_baseline(sample_solves=256, threads=None, ticks=3):
    """The oracle's C restatement of the reference tick ("port") timed on this box's host cores on a
    bounded sample of the same workload.  bench.py executes oracle/ ONLY here and in --impl reference."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cbaseline
    return cbaseline.time_baseline(sample_solves=sample_solves, threads=threads, ticks=ticks)


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle C port; the reference itself cannot be
    built here) on all host cores; one step = one tick of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cbaseline
    K, Wm = args.steps, args.warmup
    cores = os.cpu_count() or 1
    n = args.batch            # the same 1024 instances per step as the GPU arm (configs[1])
    runner = cbaseline.BaselineRunner(n, cores)
    for _ in range(Wm):
        runner.tick()
    dt = sum(runner.tick() for _ in range(K))
    value = n * K / dt
    cpu = {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": f"{K} ticks x " + runner.describe()}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batch of %d independent MPC solves per step, reference horizon, same synthetic "
                                   "instances and perturbation model as the GPU arm; CPU restatement of the reference tick "
                                   "(OSQP-style ADMM), OpenMP over instances on all host cores" % n,
                       "instances_per_gpu": n, "instances_per_step": n},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--solver", type=int, default=0)
    ap.add_argument("--horizon", default="", help="nIter,nIterSmall,controlHorizon of a horizon variant (configs[3]); "
                                                  "default: the reference horizon")
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--rollout-ticks", type=int, default=40, help="ticks of the device-resident closed-loop leg (0: skip)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-instance latency leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs only)")
    ap.add_argument("--no-extras", action="store_true", help="skip the gather / long-horizon / Monte Carlo legs")
    ap.add_argument("--mc-instances", type=int, default=65536, help="Monte Carlo sweep: instances in total over all ranks")
    ap.add_argument("--mc-ticks", type=int, default=200)
    ap.add_argument("--sweep-instances", type=int, default=2048,
                    help="parameter sweep with per-instance joint-limit rows (configs[4]): instances per GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
