"""Device-resident closed-loop rollouts of a batch of MPC instances (BASELINE.json configs[2]).

``BatchedRollout`` drives the loop body of the reference driver (src/variable_sampling_mpc.py:106-161) for B
instances without a host round trip: surrogate plant -> pack -> linearise kernel -> QP kernel -> feedback, all
on the GPU (csrc/vsmpc_plant.cu; the plant is a SURROGATE for MuJoCo, see include/vsmpc.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .batched import BatchedVSMPC, VsmpcError
from .pack import DEFAULT_JOINT_SELECTOR
from .synthetic import SyntheticRobot


def plant_model_from_robot(rb: SyntheticRobot, sel=None, dt_sim: float = 0.001, n_sub: int = 5) -> L.VsmpcPlantModel:
    """Frozen body-frame kinematics of the (synthetic) robot for the surrogate plant."""
    sel = list(DEFAULT_JOINT_SELECTOR if sel is None else sel)
    m = L.VsmpcPlantModel()
    m.com_from_base_body = (C.c_double * 3)(*rb.com_from_base_body)
    m.jet_pos_body = (C.c_double * 12)(*rb.jet_pos_body.reshape(-1))
    m.jet_axes_body = (C.c_double * 12)(*rb.jet_axes_body.reshape(-1))
    m.J_rel_ang_body = (C.c_double * 96)(*rb.J_rel_body[:, 3:6, :][:, :, sel].reshape(-1))
    m.J_jet_lin_body = (C.c_double * 96)(*rb.J_jet_lin_body[:, :, sel].reshape(-1))
    m.J_com_body = (C.c_double * 24)(*rb.J_com_body[:, sel].reshape(-1))
    m.gravity = (C.c_double * 3)(*rb.gravity)
    m.q0 = (C.c_double * 8)(*rb.joint_pos0[sel])
    m.dt_sim = float(dt_sim)
    m.n_sub = int(n_sub)
    return m


def plant_state_from_states(state: dict, sel=None) -> np.ndarray:
    """(PLANT_STATE_DOUBLES, B) SoA plant state from a getter-level batch (synthetic.make_states)."""
    sel = list(DEFAULT_JOINT_SELECTOR if sel is None else sel)
    B = state["wRb"].shape[0]
    ps = np.zeros((L.PLANT_STATE_DOUBLES, B))
    ps[L.PS_P_COM:L.PS_P_COM + 3] = state["p_com"].T
    ps[L.PS_LIN_MOM_WORLD:L.PS_LIN_MOM_WORLD + 3] = np.einsum("bij,bj->bi", state["wRb"], state["momentum_body"][:, :3]).T
    ps[L.PS_RPY:L.PS_RPY + 3] = state["rpy"].T
    ps[L.PS_ANG_MOM_BODY:L.PS_ANG_MOM_BODY + 3] = state["momentum_body"][:, 3:].T
    ps[L.PS_THRUST:L.PS_THRUST + 4] = state["thrust"].T
    ps[L.PS_THRUST_DOT:L.PS_THRUST_DOT + 4] = state["thrust_dot_est"].T
    ps[L.PS_THROTTLE:L.PS_THROTTLE + 4] = state["throttle_prev"].T
    ps[L.PS_THRUST_DES:L.PS_THRUST_DES + 4] = state["thrust_des"].T
    ps[L.PS_THRUST_DOT_DES:L.PS_THRUST_DOT_DES + 4] = state["thrust_dot_des"].T
    ps[L.PS_Q_CMD:L.PS_Q_CMD + 8] = state["q_cmd"][:, sel].T
    # jet-NN mode: the network's thrust state starts at the measured thrust, the EKF covariance at 0.1 I
    # (ironcub_mujoco_simulator.py:53-57)
    ps[L.PS_THRUST_NN:L.PS_THRUST_NN + 4] = state["thrust"].T.astype(np.float32)
    ps[L.PS_EKF_P:L.PS_EKF_P + 16] = np.tile(np.array([0.1, 0.0, 0.0, 0.1]), 4)[:, None]
    return ps


def plant_params(B: int, rb: SyntheticRobot, mass_scale=None, inertia_scale=None, thrust_disturbance=None) -> np.ndarray:
    pp = np.zeros((L.PLANT_PARAM_DOUBLES, B))
    ms = np.ones(B) if mass_scale is None else np.asarray(mass_scale, float)
    isc = np.ones(B) if inertia_scale is None else np.asarray(inertia_scale, float)
    pp[L.PP_MASS] = np.float32(rb.mass * ms).astype(np.float64)      # Robot::m_totalMass is a float
    pp[L.PP_INERTIA_BODY:L.PP_INERTIA_BODY + 9] = rb.I_body.reshape(9, 1) * isc[None, :]
    if thrust_disturbance is not None:
        pp[L.PP_THRUST_DISTURBANCE:L.PP_THRUST_DISTURBANCE + 4] = np.asarray(thrust_disturbance, float).T
    return pp


class BatchedRollout:
    """B closed loops on one GPU.  ``mpc`` is a fresh (not yet configured) BatchedVSMPC."""

    def __init__(self, mpc: BatchedVSMPC, robot: SyntheticRobot | None = None, dt_sim: float = 0.001, n_sub: int = 5):
        self.mpc = mpc
        self.robot = robot or SyntheticRobot()
        self.model = plant_model_from_robot(self.robot, mpc.sel, dt_sim, n_sub)
        self._lib = mpc._lib

    def init(self, state0: dict, mass_scale=None, inertia_scale=None, thrust_disturbance=None, phase0=None):
        """Start every loop at ``state0``; runs IMPCProblem::configure on the pack built from it."""
        B = self.mpc.B
        ps = np.ascontiguousarray(plant_state_from_states(state0, self.mpc.sel))
        pp = np.ascontiguousarray(plant_params(B, self.robot, mass_scale, inertia_scale, thrust_disturbance))
        jp = np.ascontiguousarray(state0["joint_pos"][:, self.mpc.sel].T)
        ph = None if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self.mpc._ck(self._lib.vsmpc_rollout_init(self.mpc._h, C.byref(self.model), ps.ctypes.data, pp.ctypes.data,
                                                  jp.ctypes.data, ph.ctypes.data if ph is not None else None),
                     "vsmpc_rollout_init")

    def set_jet_nn(self, weights: dict | None, ekf_R=None, ekf_Q=None):
        """Jet plant + estimator of the reference simulator: ``weights`` = dict(w_ih (320,2), b_ih, b_hh (320,), fc_w (80,),
        fc_b (1,), norm (4,)) — the numerical content of the reference's ``jet_model_torch/model_7.pth`` — or None to go
        back to the second-order jet model.  Defaults of R, Q: ironcub_mujoco_simulator.py:54-56."""
        if weights is None:
            self.mpc._ck(self._lib.vsmpc_rollout_set_jet_nn(self.mpc._h, *([None] * 8)), "vsmpc_rollout_set_jet_nn")
            return
        f32 = lambda a, n: np.ascontiguousarray(np.asarray(a, np.float32).reshape(n))
        w_ih, b_ih, b_hh = f32(weights["w_ih"], 640), f32(weights["b_ih"], 320), f32(weights["b_hh"], 320)
        fc_w, fc_b = f32(weights["fc_w"], 80), f32(weights["fc_b"], 1)
        norm = np.ascontiguousarray(np.asarray(weights["norm"], np.float64).reshape(4))
        R = np.ascontiguousarray(np.asarray(np.eye(2) * 0.5 if ekf_R is None else ekf_R, np.float64).reshape(4))
        Q = np.ascontiguousarray(np.asarray(np.eye(2) * 0.1 if ekf_Q is None else ekf_Q, np.float64).reshape(4))
        self._keep = (w_ih, b_ih, b_hh, fc_w, fc_b, norm, R, Q)
        self.mpc._ck(self._lib.vsmpc_rollout_set_jet_nn(self.mpc._h, *[a.ctypes.data for a in self._keep]),
                     "vsmpc_rollout_set_jet_nn")

    def jet_nn_eval(self, T, throttle, dt: float = 0.001):
        """One step of the neural jet plant on the device for (n, 4) float32 thrusts / throttles (parity seam)."""
        T = np.ascontiguousarray(T, dtype=np.float32); u = np.ascontiguousarray(throttle, dtype=np.float32)
        To, Td = np.empty_like(T), np.empty_like(T)
        self.mpc._ck(self._lib.vsmpc_jet_nn_eval(self.mpc._h, T.shape[0], float(dt), T.ctypes.data, u.ctypes.data,
                                                 To.ctypes.data, Td.ctypes.data), "vsmpc_jet_nn_eval")
        return To, Td

    def run(self, n_ticks: int, record_every: int = 0, use_graph: bool = True):
        """n_ticks controller ticks; returns the record array (n_rec, B, 16) or None."""
        rec = None
        if record_every > 0:
            rec = np.empty((n_ticks // record_every, self.mpc.B, L.ROLLOUT_REC_DOUBLES))
        self.mpc._ck(self._lib.vsmpc_rollout_run(self.mpc._h, int(n_ticks), int(record_every),
                                                 rec.ctypes.data if rec is not None else None, 1 if use_graph else 0),
                     "vsmpc_rollout_run")
        return rec

    def plant_state(self) -> np.ndarray:
        ps = np.empty((L.PLANT_STATE_DOUBLES, self.mpc.B))
        self.mpc._ck(self._lib.vsmpc_rollout_get_state(self.mpc._h, ps.ctypes.data), "vsmpc_rollout_get_state")
        return ps

    def pack(self) -> np.ndarray:
        pk = np.empty((L.PACK_DOUBLES, self.mpc.B))
        self.mpc._ck(self._lib.vsmpc_rollout_get_pack(self.mpc._h, pk.ctypes.data), "vsmpc_rollout_get_pack")
        return pk


LOG_KEYS = ("CoMPosition", "CoMPosition_desired", "base_orientation_desired", "base_position", "base_orientation",
            "base_lin_vel", "base_ang_vel", "base_lin_vel_filtered", "base_ang_vel_filtered", "joints_pos_meas",
            "joints_pos_ref", "linear_momentum", "angular_momentum", "momentum_reference", "estimated_thrust",
            "estimated_thrust_dot", "thrust_desired", "thrust_desired_dot", "alpha_gravity", "throttle", "mom_dot",
            "time_controller", "time_MPC")      # the 23 series of src/variable_sampling_mpc.py:163-186


def run_logged(loop: BatchedRollout, n_ticks: int, instance: int = 0, period_mpc: float = 0.005) -> dict:
    """Advance the device-resident loops tick by tick and log ONE instance with the reference driver's 23 series
    (src/variable_sampling_mpc.py:139-161 appends, :163-186 keys; same moment of the tick: measured state and the
    outputs of the tick it produced, before the plant steps on).  A logging run reads the device state back every tick
    (plant pack, output row, published references) — meant for inspecting one loop, not for throughput.
    What the surrogate plant cannot supply is said here: it has no velocity filter (``*_filtered`` = the raw values) and
    its base velocity is the CoM velocity minus omega x (R com_from_base); joints are position-controlled, so the measured
    joint vector is the commanded one (non-controlled joints at the posture of ``SyntheticRobot``)."""
    import time
    from .pack import PACK_OFFSETS as PO
    mpc, rb = loop.mpc, loop.robot
    sel = list(mpc.sel)
    i = int(instance)
    f = lambda pk, name: pk[PO[name][0]:PO[name][0] + PO[name][1], i].copy()
    log = {k: [] for k in LOG_KEYS}
    # initial rows of the series the driver seeds before its loop (:73-92)
    pk = loop.pack()
    refs = mpc.get_references()
    log["CoMPosition"].append(f(pk, "p_com"))
    log["CoMPosition_desired"].append(refs["posCoMReference"][i])
    log["base_orientation_desired"].append(refs["RPYReference"][i])
    log["linear_momentum"].append(f(pk, "momentum_body")[:3])
    log["angular_momentum"].append(f(pk, "momentum_body")[3:])
    log["momentum_reference"].append(refs["momentumReference"][i])
    q_full = np.array(rb.joint_pos0, float)
    for k in range(int(n_ticks)):
        t0 = time.perf_counter()
        loop.run(1, use_graph=False)
        out, status = mpc.get_output()
        log["time_MPC"].append(time.perf_counter() - t0)
        pk = loop.pack()
        refs = mpc.get_references()
        R = f(pk, "wRb").reshape(3, 3)
        mass = f(pk, "mass")[0]
        mom = f(pk, "momentum_body")
        w = f(pk, "omega_world")
        v_com = R @ mom[:3] / mass
        v_base = v_com - np.cross(w, f(pk, "p_com") - f(pk, "base_pos"))
        thrust = f(pk, "thrust")
        A_body = f(pk, "A_mom_body").reshape(6, 4)
        A_world = np.vstack([R @ A_body[:3], R @ A_body[3:]])          # Robot::getMatrixAmomJets() (world axes)
        qm, qr = q_full.copy(), q_full.copy()
        qm[sel] = f(pk, "q_cmd")
        qr[sel] = out[i, L.OUT_JOINTS_REF:L.OUT_JOINTS_REF + 8]
        log["CoMPosition"].append(f(pk, "p_com"))
        log["CoMPosition_desired"].append(refs["posCoMReference"][i])
        log["estimated_thrust"].append(thrust)
        log["estimated_thrust_dot"].append(f(pk, "thrust_dot_est"))
        log["thrust_desired"].append(out[i, L.OUT_THRUST:L.OUT_THRUST + 4].copy())
        log["thrust_desired_dot"].append(out[i, L.OUT_THRUST_DOT:L.OUT_THRUST_DOT + 4].copy())
        log["base_position"].append(f(pk, "base_pos"))
        log["base_orientation"].append(f(pk, "rpy"))
        log["base_orientation_desired"].append(refs["RPYReference"][i])
        log["base_lin_vel"].append(v_base)
        log["base_ang_vel"].append(w)
        log["base_lin_vel_filtered"].append(v_base)
        log["base_ang_vel_filtered"].append(w)
        log["linear_momentum"].append(mom[:3])
        log["angular_momentum"].append(mom[3:])
        log["momentum_reference"].append(refs["momentumReference"][i])
        log["alpha_gravity"].append(float(refs["alphaGravity"][i]))
        log["joints_pos_meas"].append(qm)
        log["joints_pos_ref"].append(qr)
        log["time_controller"].append(period_mpc * (k + 1))
        log["throttle"].append(out[i, L.OUT_THROTTLE:L.OUT_THROTTLE + 4].copy())
        log["mom_dot"].append(A_world @ thrust)
    data = {k: np.asarray(v, float) for k, v in log.items()}
    data["qp_status"] = np.asarray(status[i])       # extra: status of the last tick
    return data


def save_log_mat(path: str, rec, instance: int = 0, period_mpc: float = 0.005, record_every: int = 1) -> dict:
    """Write a log in the layout of the reference driver's end-of-run file (``scipy.io.savemat`` of the dictionary of
    src/variable_sampling_mpc.py:163-194).  ``rec``: the dict of ``run_logged`` (all 23 series) or the (n_rec, B, 16) record
    array of ``BatchedRollout.run`` (the six series the device loop records by itself)."""
    import scipy.io
    if isinstance(rec, dict):
        missing = [k for k in LOG_KEYS if k not in rec]
        if missing:
            raise ValueError(f"log dictionary lacks {missing}")
        scipy.io.savemat(path, rec)
        return rec
    r = np.asarray(rec)[:, instance, :]
    data = {
        "CoMPosition": r[:, 0:3],
        "base_orientation": r[:, 3:6],
        "estimated_thrust": r[:, 6:10],
        "throttle": r[:, 10:14],
        "qp_status": r[:, 14],
        "time_controller": period_mpc * record_every * (1 + np.arange(r.shape[0])),
    }
    scipy.io.savemat(path, data)
    return data
