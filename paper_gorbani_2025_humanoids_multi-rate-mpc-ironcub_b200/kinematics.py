"""Batched reduced kinematics: the kinematic tree as arrays, the robot-state SoA, and the host side of
``vsmpc_set_kin_model`` / ``vsmpc_configure_kinematics`` / ``vsmpc_set_state_kinematics`` (csrc/vsmpc_kinematics.cu).

The reference builds the kinematic rows of the pack on the host through iDynTree (``Robot::setState``,
UT/src/Robot.cpp:212-332).  Here a tree given as arrays is evaluated on the device for the whole batch, so a controller
sends the robot state (18 + 2 n_dof + 28 doubles per instance) instead of the 359-double pack.

The iRonCub URDF is not vendored with the reference: ``synthetic_humanoid()`` is a SYNTHETIC 24-link tree with the joint
order of ``src/config/robot.toml:3-27`` (3 torso joints, 4 + 4 arm joints = the controlled ones, 6 + 6 leg joints), base
frame ``chest``, jets 0 / 1 on the forearms and 2 / 3 on the chest.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

MAX_LINKS = 32
KS_WRB, KS_BASE_POS, KS_BASE_LIN_VEL, KS_OMEGA_WORLD, KS_Q = 0, 9, 12, 15, 18
PASS_FIELDS = ("thrust", "thrust_dot_est", "thrust_des", "thrust_dot_des", "throttle_prev")   # then q_cmd[sel]


class VsmpcKinModel(C.Structure):
    _fields_ = [
        ("n_links", C.c_int), ("n_dof", C.c_int), ("parent", C.c_int * MAX_LINKS), ("dof", C.c_int * MAX_LINKS),
        ("R0", (C.c_double * 9) * MAX_LINKS), ("p0", (C.c_double * 3) * MAX_LINKS), ("axis", (C.c_double * 3) * MAX_LINKS),
        ("mass", C.c_double * MAX_LINKS), ("com", (C.c_double * 3) * MAX_LINKS), ("inertia", (C.c_double * 9) * MAX_LINKS),
        ("jet_link", C.c_int * 4), ("jet_pos", (C.c_double * 3) * 4), ("jet_axis", (C.c_double * 3) * 4),
        ("delta_com", C.c_double * 3), ("gravity", C.c_double * 3), ("sel", C.c_int * 8),
    ]


def _rot(axis, ang):
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def synthetic_humanoid(seed: int = 11) -> dict:
    """A seeded synthetic humanoid: dict of arrays (the layout of ``vsmpc_kin_model`` and of oracle/kinematics_oracle.py)."""
    g = np.random.default_rng(seed)
    parent, dof, p0, axis, mass, com = [-1], [-1], [np.zeros(3)], [np.array([0.0, 0.0, 1.0])], [18.0], [np.array([0.0, 0.0, 0.05])]

    def chain(start_parent, dofs, offsets, axes, masses):
        par = start_parent
        for j, o, a, m in zip(dofs, offsets, axes, masses):
            parent.append(par)
            dof.append(j)
            p0.append(np.asarray(o, float))
            axis.append(np.asarray(a, float) / np.linalg.norm(a))
            mass.append(m)
            com.append(np.asarray(o, float) * 0.0 + g.uniform(-0.03, 0.03, 3) + np.array([0.0, 0.0, -0.05]))
            par = len(parent) - 1
        return par

    # chest (base) -> torso yaw / roll / pitch -> pelvis
    pelvis = chain(0, [2, 1, 0], [[0, 0, -0.12], [0, 0, -0.04], [0, 0, -0.04]], [[0, 0, 1], [1, 0, 0], [0, 1, 0]], [1.0, 1.0, 9.0])
    # arms: shoulder pitch / roll / yaw, elbow
    chain(0, [3, 4, 5, 6], [[0.0, 0.11, 0.10], [0.0, 0.04, 0.0], [0.0, 0.02, -0.06], [0.0, 0.0, -0.16]],
          [[0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0.2]], [1.2, 0.8, 1.6, 2.4])
    chain(0, [7, 8, 9, 10], [[0.0, -0.11, 0.10], [0.0, -0.04, 0.0], [0.0, -0.02, -0.06], [0.0, 0.0, -0.16]],
          [[0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, -0.2]], [1.2, 0.8, 1.6, 2.4])
    # legs: hip pitch / roll / yaw, knee, ankle pitch / roll
    for side, d0 in ((1.0, 11), (-1.0, 17)):
        chain(pelvis, [d0 + k for k in range(6)],
              [[0, 0.07 * side, -0.05], [0, 0.01 * side, 0], [0, 0, -0.05], [0, 0, -0.22], [0, 0, -0.24], [0, 0, -0.02]],
              [[0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0]], [1.5, 1.0, 3.0, 2.5, 0.6, 1.0])
    n = len(parent)
    R0 = [np.eye(3)] + [_rot(g.normal(size=3), g.uniform(-0.3, 0.3)) for _ in range(n - 1)]
    inertia = []
    for l in range(n):
        A = g.uniform(-1, 1, (3, 3))
        Q, _ = np.linalg.qr(A)
        inertia.append(Q @ np.diag(mass[l] * g.uniform(0.002, 0.02, 3)) @ Q.T)
    return dict(
        n_links=n, n_dof=23, parent=parent, dof=dof, R0=R0, p0=p0, axis=axis, mass=mass, com=com, inertia=inertia,
        jet_link=[7, 11, 0, 0],   # forearms (links 7 and 11), chest
        jet_pos=[np.array([0.02, 0.03, -0.10]), np.array([0.02, -0.03, -0.10]), np.array([-0.16, 0.11, 0.20]), np.array([-0.16, -0.11, 0.20])],
        jet_axis=[v / np.linalg.norm(v) for v in (np.array([0.05, 0.10, 1.0]), np.array([0.05, -0.10, 1.0]),
                                                  np.array([-0.08, 0.02, 1.0]), np.array([-0.08, -0.02, 1.0]))],
        delta_com=np.array([0.005, 0.0, -0.01]), gravity=np.array([0.0, 0.0, -9.81]), sel=list(range(3, 11)))


def model_struct(model: dict) -> VsmpcKinModel:
    n = int(model["n_links"])
    if n > MAX_LINKS:
        raise ValueError("at most 32 links")
    m = VsmpcKinModel()
    m.n_links, m.n_dof = n, int(model["n_dof"])
    for l in range(n):
        m.parent[l], m.dof[l] = int(model["parent"][l]), int(model["dof"][l])
        m.R0[l] = (C.c_double * 9)(*np.asarray(model["R0"][l], float).reshape(9))
        m.p0[l] = (C.c_double * 3)(*np.asarray(model["p0"][l], float))
        m.axis[l] = (C.c_double * 3)(*np.asarray(model["axis"][l], float))
        m.mass[l] = float(model["mass"][l])
        m.com[l] = (C.c_double * 3)(*np.asarray(model["com"][l], float))
        m.inertia[l] = (C.c_double * 9)(*np.asarray(model["inertia"][l], float).reshape(9))
    for i in range(4):
        m.jet_link[i] = int(model["jet_link"][i])
        m.jet_pos[i] = (C.c_double * 3)(*np.asarray(model["jet_pos"][i], float))
        m.jet_axis[i] = (C.c_double * 3)(*np.asarray(model["jet_axis"][i], float))
    m.delta_com = (C.c_double * 3)(*np.asarray(model["delta_com"], float))
    m.gravity = (C.c_double * 3)(*np.asarray(model["gravity"], float))
    m.sel = (C.c_int * 8)(*[int(x) for x in model["sel"]])
    return m


def kin_state_rows(n_dof: int) -> int:
    return KS_Q + 2 * n_dof + 28


def build_kin_state(model: dict, rs: dict) -> np.ndarray:
    """Robot state of a batch -> SoA (rows, B).  ``rs``: wRb (B,3,3), base_pos (B,3), base_lin_vel (B,3), omega_world (B,3),
    q (B,n_dof), qd (B,n_dof), thrust / thrust_dot_est / thrust_des / thrust_dot_des / throttle_prev (B,4), q_cmd (B,n_dof)."""
    nd = int(model["n_dof"])
    B = rs["wRb"].shape[0]
    ks = np.empty((kin_state_rows(nd), B))
    ks[KS_WRB:KS_WRB + 9] = np.asarray(rs["wRb"], float).reshape(B, 9).T
    ks[KS_BASE_POS:KS_BASE_POS + 3] = np.asarray(rs["base_pos"], float).T
    ks[KS_BASE_LIN_VEL:KS_BASE_LIN_VEL + 3] = np.asarray(rs["base_lin_vel"], float).T
    ks[KS_OMEGA_WORLD:KS_OMEGA_WORLD + 3] = np.asarray(rs["omega_world"], float).T
    ks[KS_Q:KS_Q + nd] = np.asarray(rs["q"], float).T
    ks[KS_Q + nd:KS_Q + 2 * nd] = np.asarray(rs["qd"], float).T
    o = KS_Q + 2 * nd
    for name in PASS_FIELDS:
        ks[o:o + 4] = np.asarray(rs[name], float).T
        o += 4
    ks[o:o + 8] = np.asarray(rs["q_cmd"], float)[:, list(model["sel"])].T
    return np.ascontiguousarray(ks)


class KinematicsFrontEnd:
    """``BatchedVSMPC`` driven by robot states: the pack is built on the device by the kinematics kernel."""

    def __init__(self, mpc, model: dict):
        self.mpc, self.model = mpc, model
        self._struct = model_struct(model)
        mpc._ck(mpc._lib.vsmpc_set_kin_model(mpc._h, C.byref(self._struct)), "vsmpc_set_kin_model")
        self.rows = mpc._lib.vsmpc_kin_state_doubles(mpc._h)
        assert self.rows == kin_state_rows(int(model["n_dof"]))

    def _ks(self, robot_state):
        ks = robot_state if isinstance(robot_state, np.ndarray) else build_kin_state(self.model, robot_state)
        if ks.shape != (self.rows, self.mpc.B) or ks.dtype != np.float64 or not ks.flags.c_contiguous:
            raise ValueError(f"robot-state SoA must be a contiguous float64 array of shape {(self.rows, self.mpc.B)}")
        return ks

    def configure(self, robot_state, phase0=None):
        ks = self._ks(robot_state)
        ph = None if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self.mpc._ck(self.mpc._lib.vsmpc_configure_kinematics(self.mpc._h, ks.ctypes.data, ph.ctypes.data if ph is not None else None),
                     "vsmpc_configure_kinematics")

    def update(self, robot_state):
        ks = self._ks(robot_state)
        self._keep = ks       # the copy is asynchronous: keep the buffer alive until the next call
        self.mpc._ck(self.mpc._lib.vsmpc_set_state_kinematics(self.mpc._h, ks.ctypes.data), "vsmpc_set_state_kinematics")

    def pack(self) -> np.ndarray:
        pk = np.empty((L.PACK_DOUBLES, self.mpc.B))
        self.mpc._ck(self.mpc._lib.vsmpc_get_kinematics_pack(self.mpc._h, pk.ctypes.data), "vsmpc_get_kinematics_pack")
        return pk
