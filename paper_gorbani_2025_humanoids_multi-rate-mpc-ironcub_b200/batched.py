"""Batched host-side mirror of the reference's ``VariableSamplingMPC`` over the C-ABI.

``BatchedVSMPC`` keeps the reference's configure / update / solveMPC / get* surface
(variableSamplingMPC.h:15-41; MPCPyBindings.cpp:22-90) with every array carrying a leading
instance dimension B.  All compute happens in libvsmpc.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .config import JET_COEFF, JET_NORM, default_params
from .pack import PACK_DOUBLES, build_pack, DEFAULT_JOINT_SELECTOR


class VsmpcError(RuntimeError):
    pass


def _cfg_struct(params: dict, traj: dict, solver: int, keep: list) -> L.VsmpcConfig:
    p = dict(default_params())
    p.update(params or {})
    c = L.VsmpcConfig()
    c.n_iter, c.n_iter_small, c.control_horizon = int(p["nIter"]), int(p["nIterSmall"]), int(p["controlHorizon"])
    c.period_mpc, c.period_large, c.period_small = p["periodMPC"], p["periodMPCLargeSteps"], p["periodMPCSmallSteps"]
    c.use_jet_dynamic = int(bool(p["useJetDynamic"]))
    c.use_estimated_thrust = int(bool(p["useEstimatedThrust"]))
    opt = p["jointsLambdaOption"]
    if opt not in ("unfiltered", "constant"):
        raise VsmpcError("Parameter 'jointsLambdaOption' should be 'unfiltered' or 'constant'.")
    c.joints_lambda_option = 0 if opt == "unfiltered" else 1
    for name, key in (("weight_com_pos", "weightCoMPos"), ("weight_com_pos_error", "weightCoMPosError"),
                      ("weight_lin_mom", "weightLinMom"), ("weight_rpy", "weightRPY"),
                      ("weight_rpy_error", "weightRPYError"), ("weight_ang_mom", "weightAngMom")):
        v = list(p[key])
        if len(v) != 3:
            raise VsmpcError(f"Parameter '{key}' must have 3 elements")
        setattr(c, name, (C.c_double * 3)(*v))
    wdj = list(p["weightDeltaJoint"])
    if len(wdj) != 8:
        raise VsmpcError("The size of the vector containing the weights for the joint deltas is not correct.")
    c.weight_delta_joint = (C.c_double * 8)(*wdj)
    c.weight_throttle = p["weightThrottle"]
    c.weight_initial_throttle = p["weightInitialThrottle"]
    c.weight_regularization_joint_pos = p["weightRegularizationJointPos"]
    c.throttle_min, c.throttle_max = p["throttleMin"], p["throttleMax"]
    c.jet_coeff = (C.c_double * 13)(*p.get("jetCoeff", JET_COEFF))
    c.jet_norm = (C.c_double * 4)(*p.get("jetNorm", JET_NORM))

    def arr(a, shape_rows):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if a.ndim != 2 or a.shape[0] != shape_rows:
            raise VsmpcError(f"trajectory array must be ({shape_rows}, n)")
        t = np.ascontiguousarray(a.T)  # sample-major
        keep.append(t)
        return t.ctypes.data_as(L.c_double_p), a.shape[1]

    c.alpha_gravity, c.alpha_len = arr(traj["alphaGravity"], 1)
    c.alpha_fps = int(traj["alpha_fps"])
    c.position_com, n = arr(traj["positionCoM"], 3)
    c.velocity_com, n2 = arr(traj["velocityCoM"], 3)
    c.rpy, n3 = arr(traj["RPY"], 3)
    c.rpy_dot, n4 = arr(traj["RPYDot"], 3)
    if not (n == n2 == n3 == n4):
        raise VsmpcError("trajectory arrays differ in length")
    c.traj_len, c.traj_fps = n, int(traj["traj_fps"])
    c.solver = int(solver)
    # optional joint-limit rows (JointPositionConstraint, constraintsVSMPC.cpp:388-468; parameters jointPos_max / jointPos_min
    # in degrees, :421-423 — absent from the shipped XML, hence off by default)
    jmax, jmin = p.get("jointPos_max"), p.get("jointPos_min")
    if (jmax is None) != (jmin is None):
        raise VsmpcError("Parameters 'jointPos_max' and 'jointPos_min' go together")
    if jmax is not None:
        if len(jmax) != 8 or len(jmin) != 8:
            raise VsmpcError("The size of the vector containing the joint position limits is not correct.")
        c.use_joint_limits = 1
        c.joint_pos_min_deg = (C.c_double * 8)(*[float(x) for x in jmin])
        c.joint_pos_max_deg = (C.c_double * 8)(*[float(x) for x in jmax])
    return c


class BatchedVSMPC:
    """B independent MPC instances on one GPU."""

    def __init__(self, n_instances: int, params: dict | None, trajectories: dict, device: int = 0,
                 solver: int = 0, full_solution: bool = False):
        self._lib = L.load()
        self.B = int(n_instances)
        self.params = dict(default_params())
        self.params.update(params or {})
        keep: list = []
        cfg = _cfg_struct(self.params, trajectories, solver, keep)
        h = C.c_void_p()
        rc = self._lib.vsmpc_create(C.byref(cfg), self.B, int(device), C.byref(h))
        self._h = h
        if rc != L.OK:
            msg = self._lib.vsmpc_last_error(h).decode() if h else "vsmpc_create failed"
            if h:
                self._lib.vsmpc_destroy(h)
            self._h = None
            raise VsmpcError(f"vsmpc_create: {msg} (code {rc})")
        self.n_var = self._lib.vsmpc_n_var(h)
        self.n_con = self._lib.vsmpc_n_constraints(h)
        self.sel = list(DEFAULT_JOINT_SELECTOR)
        self.device = int(device)
        self._ck(self._lib.vsmpc_set_full_solution(h, 1 if full_solution else 0), "vsmpc_set_full_solution")

    # ---- lifecycle -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.vsmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != L.OK:
            raise VsmpcError(f"{what}: {self._lib.vsmpc_last_error(self._h).decode()} (code {rc})")

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self._lib.vsmpc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "vsmpc_set_stream")

    @staticmethod
    def _f64(a, shape):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != shape:
            raise VsmpcError(f"expected array of shape {shape}, got {a.shape}")
        return a

    def set_instance_params(self, jet_coeff=None, jet_norm=None, throttle_min=None, throttle_max=None):
        """Per-instance jet model (B,13)/(B,4) and throttle limits (B,) — configs[4] parameter sweeps.  ``None``
        fields keep the handle-wide value; all ``None`` switches the per-instance table off."""
        if jet_coeff is None and jet_norm is None and throttle_min is None and throttle_max is None:
            self._ck(self._lib.vsmpc_set_instance_params(self._h, None), "vsmpc_set_instance_params")
            return
        B = self.B
        ip = np.empty((L.INSTANCE_PARAM_DOUBLES, B))
        ip[L.IP_JET_COEFF:L.IP_JET_COEFF + 13] = np.asarray(self.params.get("jetCoeff", JET_COEFF), float)[:, None] if jet_coeff is None else self._f64(jet_coeff, (B, 13)).T
        ip[L.IP_JET_NORM:L.IP_JET_NORM + 4] = np.asarray(self.params.get("jetNorm", JET_NORM), float)[:, None] if jet_norm is None else self._f64(jet_norm, (B, 4)).T
        ip[L.IP_THROTTLE_MIN] = self.params["throttleMin"] if throttle_min is None else self._f64(throttle_min, (B,))
        ip[L.IP_THROTTLE_MAX] = self.params["throttleMax"] if throttle_max is None else self._f64(throttle_max, (B,))
        ip = np.ascontiguousarray(ip)
        self._ck(self._lib.vsmpc_set_instance_params(self._h, ip.ctypes.data), "vsmpc_set_instance_params")

    def set_joint_limits(self, q_min=None, q_max=None):
        """Per-instance joint limits [rad] of the controlled joints, (B, 8) each — configs[4] "per-instance constraint
        sets"; needs jointPos_max / jointPos_min in the parameters.  ``None, None``: back to the handle-wide limits."""
        if q_min is None and q_max is None:
            self._ck(self._lib.vsmpc_set_joint_limits(self._h, None, None), "vsmpc_set_joint_limits")
            return
        lo = np.ascontiguousarray(self._f64(q_min, (self.B, 8)).T)
        hi = np.ascontiguousarray(self._f64(q_max, (self.B, 8)).T)
        self._ck(self._lib.vsmpc_set_joint_limits(self._h, lo.ctypes.data, hi.ctypes.data), "vsmpc_set_joint_limits")

    # ---- reference surface -------------------------------------------------------------------------
    def configure_pack(self, pack: np.ndarray, joint_pos_sel: np.ndarray, phase0=None) -> bool:
        """IMPCProblem::configure with the SoA pack (PACK_DOUBLES, B) and joint_pos_sel (8, B)."""
        pack = self._f64(pack, (PACK_DOUBLES, self.B))
        jp = self._f64(joint_pos_sel, (L.NJ, self.B))
        ph = None
        if phase0 is not None:
            ph = np.ascontiguousarray(phase0, dtype=np.int32)
            if ph.shape != (self.B,):
                raise VsmpcError("phase0 must have shape (B,)")
        self._ck(self._lib.vsmpc_configure(self._h, pack.ctypes.data, jp.ctypes.data,
                                           ph.ctypes.data if ph is not None else None), "vsmpc_configure")
        return True

    def configure(self, state: dict, phase0=None) -> bool:
        """``state``: getter-level batch dict (see synthetic.make_states)."""
        pack = build_pack(state, self.sel)
        jp = np.ascontiguousarray(state["joint_pos"][:, self.sel].T)
        return self.configure_pack(pack, jp, phase0)

    def update_pack(self, pack: np.ndarray) -> bool:
        pack = self._f64(pack, (PACK_DOUBLES, self.B))
        self._ck(self._lib.vsmpc_set_state(self._h, pack.ctypes.data), "vsmpc_set_state")
        return True

    def update_ptr(self, host_ptr: int) -> bool:
        """update from a caller-owned (e.g. pinned) host buffer of PACK_DOUBLES*B doubles."""
        self._ck(self._lib.vsmpc_set_state(self._h, C.c_void_p(host_ptr)), "vsmpc_set_state")
        return True

    def update_device_ptr(self, dev_ptr: int) -> bool:
        self._ck(self._lib.vsmpc_set_state_device(self._h, C.c_void_p(dev_ptr)), "vsmpc_set_state_device")
        return True

    def update(self, state: dict) -> bool:
        return self.update_pack(build_pack(state, self.sel))

    def linearise(self, state: dict) -> bool:
        """K1 alone (SURVEY §8b seam ii): pack + H2D + linearise kernel; read with get_dynamics / get_qp_vectors."""
        pack = self._f64(build_pack(state, self.sel), (PACK_DOUBLES, self.B))
        self._ck(self._lib.vsmpc_linearise(self._h, pack.ctypes.data), "vsmpc_linearise")
        return True

    def solve_qp(self) -> bool:
        """K2 alone (SURVEY §8b seam iii): solve the QP the last linearise left on the device."""
        self._ck(self._lib.vsmpc_solve_qp(self._h), "vsmpc_solve_qp")
        return True

    def solveMPC(self) -> bool:
        self._ck(self._lib.vsmpc_solve(self._h), "vsmpc_solve")
        return True

    def solve_async(self):
        self._ck(self._lib.vsmpc_solve_async(self._h), "vsmpc_solve_async")

    def wait(self):
        self._ck(self._lib.vsmpc_wait(self._h), "vsmpc_wait")

    # ---- getters ---------------------------------------------------------------------------------------
    def get_output(self):
        out = np.empty((self.B, L.OUT_DOUBLES))
        status = np.empty(self.B, dtype=np.int32)
        self._ck(self._lib.vsmpc_get_output(self._h, out.ctypes.data, status.ctypes.data), "vsmpc_get_output")
        return out, status

    def get_output_into(self, out_ptr: int, status_ptr: int):
        self._ck(self._lib.vsmpc_get_output(self._h, C.c_void_p(out_ptr), C.c_void_p(status_ptr)), "vsmpc_get_output")

    def get_output_async(self, out_ptr: int, status_ptr: int) -> int:
        """Enqueue the D2H copies behind the solve; returns the ticket for ``wait_output``."""
        t = C.c_int(-1)
        self._ck(self._lib.vsmpc_get_output_async(self._h, C.c_void_p(out_ptr), C.c_void_p(status_ptr), C.byref(t)),
                 "vsmpc_get_output_async")
        return t.value

    def wait_output(self, ticket: int):
        self._ck(self._lib.vsmpc_wait_output(self._h, int(ticket)), "vsmpc_wait_output")

    def output_device_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self._lib.vsmpc_get_output_device(self._h, C.byref(a), C.byref(b)), "vsmpc_get_output_device")
        return a.value, b.value

    def getSolution(self) -> np.ndarray:
        z = np.empty((self.B, self.n_var))
        self._ck(self._lib.vsmpc_get_full_solution(self._h, z.ctypes.data), "vsmpc_get_full_solution")
        return z

    def getThrottleReference(self): return self.get_output()[0][:, L.OUT_THROTTLE:L.OUT_THROTTLE + 4]
    def getThrustReference(self): return self.get_output()[0][:, L.OUT_THRUST:L.OUT_THRUST + 4]
    def getThrustDotReference(self): return self.get_output()[0][:, L.OUT_THRUST_DOT:L.OUT_THRUST_DOT + 4]
    def getDeltaJoints(self): return self.get_output()[0][:, L.OUT_DELTA_Q:L.OUT_DELTA_Q + 8]
    def getJointsReferencePositionControlled(self): return self.get_output()[0][:, L.OUT_JOINTS_REF:L.OUT_JOINTS_REF + 8]
    def getFinalState(self): return self.get_output()[0][:, L.OUT_FINAL_STATE:L.OUT_FINAL_STATE + 26]
    def getQPProblemStatus(self): return self.get_output()[1]
    def getNOptimizationVariables(self): return self.n_var
    def getNConstraints(self): return self.n_con

    # ---- inner seams (parity tests) -----------------------------------------------------------------------
    def get_dynamics(self):
        B = self.B
        A = np.empty((B, 26, 26)); BJ = np.empty((B, 26, 8)); BT = np.empty((B, 26, 4)); c = np.empty((B, 26))
        dt = np.empty(int(self.params["nIter"]))
        self._ck(self._lib.vsmpc_get_dynamics(self._h, A.ctypes.data, BJ.ctypes.data, BT.ctypes.data, c.ctypes.data,
                                              dt.ctypes.data), "vsmpc_get_dynamics")
        return A, BJ, BT, c, dt

    def get_qp_vectors(self):
        q = np.empty((self.B, self.n_var)); l = np.empty((self.B, self.n_con)); u = np.empty((self.B, self.n_con))
        self._ck(self._lib.vsmpc_get_qp_vectors(self._h, q.ctypes.data, l.ctypes.data, u.ctypes.data),
                 "vsmpc_get_qp_vectors")
        return q, l, u

    def get_counts(self):
        nf = np.empty(self.B, dtype=np.int32); ns = np.empty(self.B, dtype=np.int32)
        self._ck(self._lib.vsmpc_get_counts(self._h, nf.ctypes.data, ns.ctypes.data), "vsmpc_get_counts")
        return nf, ns

    def get_pivot_counts(self):
        """Exchange pivots of the reduced throttle QP per instance in the last solve (inverse + active set)."""
        n = np.empty(self.B, dtype=np.int32)
        self._ck(self._lib.vsmpc_get_pivot_counts(self._h, n.ctypes.data), "vsmpc_get_pivot_counts")
        return n

    def get_references(self) -> dict:
        """The QPInput fields the path writes (costsVSMPC.cpp:155-160, systemDynamicsVSMPC.cpp:310), per instance:
        alphaGravity (B,), posCoMReference (B,3), RPYReference (B,3), momentumReference (B,6)."""
        r = np.empty((self.B, L.REF_DOUBLES))
        self._ck(self._lib.vsmpc_get_references(self._h, r.ctypes.data), "vsmpc_get_references")
        return dict(alphaGravity=r[:, L.REF_ALPHA_GRAVITY].copy(), posCoMReference=r[:, L.REF_POS_COM:L.REF_POS_COM + 3].copy(),
                    RPYReference=r[:, L.REF_RPY:L.REF_RPY + 3].copy(),
                    momentumReference=r[:, L.REF_MOMENTUM:L.REF_MOMENTUM + 6].copy())

    def getHessian(self, instance: int = 0) -> np.ndarray:
        """IMPCProblem::getHessian of one instance, dense (n_var, n_var)."""
        P = np.empty((self.n_var, self.n_var))
        self._ck(self._lib.vsmpc_get_hessian(self._h, int(instance), P.ctypes.data), "vsmpc_get_hessian")
        return P

    def getLinearConstraintMatrix(self, instance: int = 0) -> np.ndarray:
        """IMPCProblem::getLinearConstraintMatrix of one instance for the current tick, dense (n_con, n_var)."""
        A = np.empty((self.n_con, self.n_var))
        self._ck(self._lib.vsmpc_get_constraint_matrix(self._h, int(instance), A.ctypes.data), "vsmpc_get_constraint_matrix")
        return A

    def set_fallback(self, mode: int):
        """Fallback QP kernel (pivoted LU of the KKT system): 0 off, 1 on (default), 2 every instance (tests)."""
        self._ck(self._lib.vsmpc_set_fallback(self._h, int(mode)), "vsmpc_set_fallback")

    def set_warm_start(self, enable: bool):
        """Long horizons: start the active set from the previous solve's working set (default) or cold."""
        self._ck(self._lib.vsmpc_set_warm_start(self._h, 1 if enable else 0), "vsmpc_set_warm_start")

    def debug_set_working_set(self, wset):
        """Tests: overwrite the stored working sets, (B, 4 * throttle blocks) of +1 / -1 / 0."""
        nblk = int(self.params["controlHorizon"]) - int(self.params["nIterSmall"]) + 1
        w = np.ascontiguousarray(wset, dtype=np.int8)
        if w.shape != (self.B, 4 * nblk):
            raise VsmpcError(f"expected working sets of shape {(self.B, 4 * nblk)}")
        self._ck(self._lib.vsmpc_debug_set_working_set(self._h, w.ctypes.data), "vsmpc_debug_set_working_set")

    def debug_set_counters(self, ref_counter: int = -1, throttle_counter: int = -1):
        self._ck(self._lib.vsmpc_debug_set_counters(self._h, ref_counter, throttle_counter), "vsmpc_debug_set_counters")


class MultiGpuVSMPC:
    """B instances sharded over several GPUs inside ONE process (C-ABI ``vsmpc_multi_*``: contiguous ranges, one handle and
    stream set per device, no inter-GPU traffic; ``devices`` may repeat an index to put several shards on one GPU).  Same
    call sequence and array layouts as ``BatchedVSMPC`` for the whole batch."""

    def __init__(self, n_instances: int, params: dict | None, trajectories: dict, n_gpus: int, devices=None, solver: int = 0,
                 full_solution: bool = False):
        self._lib = L.load()
        self.B, self.n_gpus = int(n_instances), int(n_gpus)
        self.params = dict(default_params())
        self.params.update(params or {})
        keep: list = []
        cfg = _cfg_struct(self.params, trajectories, solver, keep)
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        if dev is not None and dev.shape != (self.n_gpus,):
            raise VsmpcError("devices must have one entry per shard")
        h = C.c_void_p()
        rc = self._lib.vsmpc_create_multi(C.byref(cfg), self.B, self.n_gpus, dev.ctypes.data if dev is not None else None,
                                          C.byref(h))
        self._h = h
        if rc != L.OK:
            msg = self._lib.vsmpc_multi_last_error(h).decode() if h else "vsmpc_create_multi failed"
            if h:
                self._lib.vsmpc_multi_destroy(h)
            self._h = None
            raise VsmpcError(f"vsmpc_create_multi: {msg} (code {rc})")
        self.sel = list(DEFAULT_JOINT_SELECTOR)
        N, Ns, Nc = int(self.params["nIter"]), int(self.params["nIterSmall"]), int(self.params["controlHorizon"])
        self.n_var = 26 * (N + 1) + 8 * Nc + 4 * (Nc - Ns + 1)
        self._ck(self._lib.vsmpc_multi_set_full_solution(h, 1 if full_solution else 0), "vsmpc_multi_set_full_solution")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vsmpc_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != L.OK:
            raise VsmpcError(f"{what}: {self._lib.vsmpc_multi_last_error(self._h).decode()} (code {rc})")

    def shards(self):
        out = []
        for g in range(self._lib.vsmpc_multi_n_shards(self._h)):
            a, b = C.c_int(), C.c_int()
            self._ck(self._lib.vsmpc_multi_shard(self._h, g, C.byref(a), C.byref(b), None), "vsmpc_multi_shard")
            out.append((a.value, b.value))
        return out

    def configure_pack(self, pack, joint_pos_sel, phase0=None):
        pack = BatchedVSMPC._f64(pack, (PACK_DOUBLES, self.B))
        jp = BatchedVSMPC._f64(joint_pos_sel, (L.NJ, self.B))
        ph = None if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self._ck(self._lib.vsmpc_multi_configure(self._h, pack.ctypes.data, jp.ctypes.data,
                                                 ph.ctypes.data if ph is not None else None), "vsmpc_multi_configure")
        return True

    def configure(self, state: dict, phase0=None):
        return self.configure_pack(build_pack(state, self.sel), np.ascontiguousarray(state["joint_pos"][:, self.sel].T), phase0)

    def update_pack(self, pack):
        pack = BatchedVSMPC._f64(pack, (PACK_DOUBLES, self.B))
        self._ck(self._lib.vsmpc_multi_set_state(self._h, pack.ctypes.data), "vsmpc_multi_set_state")
        return True

    def update(self, state: dict):
        return self.update_pack(build_pack(state, self.sel))

    def solveMPC(self):
        self._ck(self._lib.vsmpc_multi_solve(self._h), "vsmpc_multi_solve")
        return True

    def get_output(self):
        out = np.empty((self.B, L.OUT_DOUBLES))
        status = np.empty(self.B, dtype=np.int32)
        self._ck(self._lib.vsmpc_multi_get_output(self._h, out.ctypes.data, status.ctypes.data), "vsmpc_multi_get_output")
        return out, status

    def getSolution(self):
        z = np.empty((self.B, self.n_var))
        self._ck(self._lib.vsmpc_multi_get_full_solution(self._h, z.ctypes.data), "vsmpc_multi_get_full_solution")
        return z
