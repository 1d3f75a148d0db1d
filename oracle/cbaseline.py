"""ctypes wrapper of oracle/c/vsmpc_ref.c (the CPU restatement of the reference tick incl. an OSQP-style
solver) + the timing loop used by bench.py's cpu_baseline / `--impl reference` legs.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import build_c

_lib = None


class RefConfig(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int), ("n_small", C.c_int), ("n_ctrl", C.c_int),
        ("period_mpc", C.c_double), ("period_large", C.c_double), ("period_small", C.c_double),
        ("use_jet_dynamic", C.c_int), ("use_estimated_thrust", C.c_int),
        ("w_com", C.c_double * 3), ("w_com_err", C.c_double * 3), ("w_lin", C.c_double * 3),
        ("w_rpy", C.c_double * 3), ("w_rpy_err", C.c_double * 3), ("w_ang", C.c_double * 3),
        ("w_dq", C.c_double * 8), ("w_throttle", C.c_double), ("w_init_throttle", C.c_double), ("w_reg_q", C.c_double),
        ("throttle_min", C.c_double), ("throttle_max", C.c_double),
        ("jc", C.c_double * 13), ("jn", C.c_double * 4),
        ("alpha", C.POINTER(C.c_double)), ("alpha_len", C.c_int),
        ("tpos", C.POINTER(C.c_double)), ("tvel", C.POINTER(C.c_double)), ("trpy", C.POINTER(C.c_double)),
        ("trpyd", C.POINTER(C.c_double)), ("traj_len", C.c_int),
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha_relax", C.c_double), ("eps_abs", C.c_double),
        ("eps_rel", C.c_double), ("delta", C.c_double),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("scaling_iters", C.c_int), ("adaptive_rho", C.c_int),
        ("adaptive_rho_interval", C.c_int), ("polish", C.c_int), ("polish_refine_iter", C.c_int),
        ("adaptive_rho_tolerance", C.c_double),
    ]


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c.build())
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.POINTER(RefConfig)]
        lib.ref_default_settings.argtypes = [C.POINTER(RefConfig)]
        for f in ("ref_destroy",):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.ref_configure.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_update.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_solve.argtypes = [C.c_void_p]
        lib.ref_get_output.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_get_solution.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_get_qp.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        for f in ("ref_nvar", "ref_ncon", "ref_iters", "ref_polished", "ref_refactors"):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.ref_tick_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def _upsample(values, fps, des_fps):
    """Trajectory::upsample (UT/src/TrajectoryManager.cpp:23-39)."""
    ratio = float(des_fps) / fps
    out = []
    for i in range(values.shape[1] - 1):
        k = 0
        while k < ratio:
            out.append(values[:, i] + (values[:, i + 1] - values[:, i]) * (k / ratio))
            k += 1
    return np.array(out).T


class RefMPC:
    """One reference-style MPC instance in C.  `traj` uses the layout of helpers.load_trajectories()."""

    def __init__(self, params: dict, traj: dict, settings: dict | None = None):
        lib = load()
        c = RefConfig()
        lib.ref_default_settings(C.byref(c))
        p = params
        c.n_iter, c.n_small, c.n_ctrl = p["nIter"], p["nIterSmall"], p["controlHorizon"]
        c.period_mpc, c.period_large, c.period_small = p["periodMPC"], p["periodMPCLargeSteps"], p["periodMPCSmallSteps"]
        c.use_jet_dynamic, c.use_estimated_thrust = int(p["useJetDynamic"]), int(p["useEstimatedThrust"])
        for name, key in (("w_com", "weightCoMPos"), ("w_com_err", "weightCoMPosError"), ("w_lin", "weightLinMom"),
                          ("w_rpy", "weightRPY"), ("w_rpy_err", "weightRPYError"), ("w_ang", "weightAngMom")):
            setattr(c, name, (C.c_double * 3)(*p[key]))
        c.w_dq = (C.c_double * 8)(*p["weightDeltaJoint"])
        c.w_throttle, c.w_init_throttle, c.w_reg_q = p["weightThrottle"], p["weightInitialThrottle"], p["weightRegularizationJointPos"]
        c.throttle_min, c.throttle_max = p["throttleMin"], p["throttleMax"]
        from .vsmpc_oracle import JetModel
        jm = JetModel()
        c.jc = (C.c_double * 13)(*jm.c)
        c.jn = (C.c_double * 4)(*jm.n)
        a = traj["TRAJECTORY_MANAGER"]
        alpha = np.asarray(a["arrays"]["alphaGravity"], float)
        des = int(1 / p["periodMPC"])
        if a["fps"] != des and alpha.shape[1] > 1:
            alpha = _upsample(alpha, a["fps"], des)
        t = traj["POSITION_TRAJECTORY"]
        arrs = {k: np.asarray(v, float) for k, v in t["arrays"].items()}
        des2 = int(1 / p["periodMPCLargeSteps"])
        if t["fps"] != des2:
            arrs = {k: _upsample(v, t["fps"], des2) for k, v in arrs.items()}
        self._keep = [np.ascontiguousarray(alpha[0])] + [np.ascontiguousarray(arrs[k].T) for k in
                                                         ("positionCoM", "velocityCoM", "RPY", "RPYDot")]
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        c.alpha, c.alpha_len = dp(self._keep[0]), len(self._keep[0])
        c.tpos, c.tvel, c.trpy, c.trpyd = [dp(x) for x in self._keep[1:]]
        c.traj_len = self._keep[1].shape[0]
        for k, v in (settings or {}).items():
            setattr(c, k, v)
        self._cfg = c
        self._lib = lib
        self.h = lib.ref_create(C.byref(c))
        self.nvar, self.ncon = lib.ref_nvar(self.h), lib.ref_ncon(self.h)

    def __del__(self):
        try:
            self._lib.ref_destroy(self.h)
        except Exception:
            pass

    def configure(self, pack_col: np.ndarray, joint_pos_sel: np.ndarray):
        pk = np.ascontiguousarray(pack_col, dtype=np.float64)
        jp = np.ascontiguousarray(joint_pos_sel, dtype=np.float64)
        self._lib.ref_configure(self.h, pk.ctypes.data, jp.ctypes.data)

    def update(self, pack_col: np.ndarray):
        pk = np.ascontiguousarray(pack_col, dtype=np.float64)
        self._lib.ref_update(self.h, pk.ctypes.data)

    def solve(self) -> int:
        return self._lib.ref_solve(self.h)

    def output(self):
        o = np.empty(54)
        self._lib.ref_get_output(self.h, o.ctypes.data)
        return o

    def solution(self):
        z = np.empty(self.nvar)
        self._lib.ref_get_solution(self.h, z.ctypes.data)
        return z

    def qp(self):
        P = np.empty((self.nvar, self.nvar)); q = np.empty(self.nvar); A = np.empty((self.ncon, self.nvar))
        l = np.empty(self.ncon); u = np.empty(self.ncon)
        self._lib.ref_get_qp(self.h, P.ctypes.data, q.ctypes.data, A.ctypes.data, l.ctypes.data, u.ctypes.data)
        return P, q, A, l, u

    def set_throttle_counter(self, v):  # test hook mirroring vsmpc_debug_set_counters
        raise NotImplementedError

    @property
    def iters(self): return self._lib.ref_iters(self.h)
    @property
    def polished(self): return bool(self._lib.ref_polished(self.h))


class BaselineRunner:
    """n reference-style instances (C), configured once on the nominal state; every tick() runs
    update + solve (warm-started) on a fresh perturbed batch with `threads` OpenMP threads."""

    def __init__(self, n: int, threads: int | None = None, seed: int = 20251002):
        import importlib
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "tests"))
        from helpers import load_trajectories
        from .vsmpc_oracle import default_params
        pkgname = "paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200"
        self.syn = importlib.import_module(pkgname + ".synthetic")
        self.pack = importlib.import_module(pkgname + ".pack")
        self.lib = load()
        self.n = int(n)
        self.threads = int(threads or os.cpu_count() or 1)
        self.seed = seed
        traj = load_trajectories()
        params = default_params()
        nom = self.syn.make_states(self.n, perturbed=False)
        nom_pack = np.ascontiguousarray(self.pack.build_pack(nom).T)
        jp = np.ascontiguousarray(nom["joint_pos"][:, self.pack.DEFAULT_JOINT_SELECTOR])
        self.inst = [RefMPC(params, traj) for _ in range(self.n)]
        for i, m in enumerate(self.inst):
            m.configure(nom_pack[i], jp[i])
        self.handles = (C.c_void_p * self.n)(*[m.h for m in self.inst])
        self.status = np.zeros(self.n, dtype=np.int32)
        self.t = 0
        # a few distinct perturbed batches, cycled
        self.packs = [np.ascontiguousarray(self.pack.build_pack(
            self.syn.make_states(self.n, seed=seed + 7919 * j, perturbed=True)).T) for j in range(4)]
        self.tick()  # untimed first tick: OSQP setup (ordering, symbolic factorisation), like the reference's first solve

    def tick(self) -> float:
        pk = self.packs[self.t % len(self.packs)]
        self.t += 1
        t0 = time.perf_counter()
        self.lib.ref_tick_batch(self.handles, self.n, pk.ctypes.data, self.threads, self.status.ctypes.data)
        return time.perf_counter() - t0

    def describe(self) -> str:
        iters = float(np.mean([m.iters for m in self.inst]))
        return (f"{self.n} instances of the bench workload per tick, warm-started; C restatement of the reference tick "
                f"(dense assembly + OSQP-style ADMM, eps 1e-3, polish; gcc -O3 -march=native; OpenMP over instances, "
                f"{self.threads} threads); mean ADMM iterations {iters:.1f}; solved {int((self.status == 1).sum())}/{self.n}")


def time_baseline(sample_solves: int = 256, threads: int | None = None, ticks: int = 3, seed: int = 20251002):
    """cpu_baseline dict of bench.py: bounded sample of the bench workload on the host cores."""
    r = BaselineRunner(sample_solves, threads, seed)
    r.tick()
    dt = sum(r.tick() for _ in range(ticks))
    return {"value": r.n * ticks / dt, "unit": "solves/s", "cores": r.threads, "kind": "port",
            "sample": f"{ticks} ticks x " + r.describe(), "ms_per_solve_per_core": 1e3 * dt * r.threads / (r.n * ticks)}
