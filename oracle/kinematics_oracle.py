"""TEST INFRASTRUCTURE — NumPy restatement of the part of ``Robot::setState`` the MPC path consumes
(UT/src/Robot.cpp:212-278,325-332), for a kinematic tree given as arrays.

What the reference computes there through iDynTree's KinDynComputations (setRobotState, getFreeFloatingMassMatrix,
getCentroidalTotalMomentum, getCenterOfMassPosition, getCenterOfMassJacobian, getFrameFreeFloatingJacobian,
getRelativeJacobian, getWorldTransform; MIXED velocity representation, iDynTree's default) is restated here link by link,
in the most literal form (per-link Jacobians summed, no composite / subtree shortcuts), so that it is independent of the
CUDA kernel's formulation (csrc/vsmpc_kinematics.cu: subtree moments, one warp per instance).

PARITY UNPINNED against iDynTree itself: iDynTree and the iRonCub URDF (ironcub-models 0.0.2) are not in this image, so the
conventions below are iDynTree's documented ones, checked here against finite differences of the forward kinematics and
closed-form chains (tests/test_kinematics.py), not against the library.  The 23-DoF tree of the tests is synthetic.

Conventions (iDynTree MIXED representation):
  * base velocity = (velocity of the base origin in world axes, angular velocity in world axes);
  * free-floating Jacobian of a frame: linear rows = velocity of the frame origin (world axes), angular rows world axes;
  * relative Jacobian getRelativeJacobian(base, frame): origin = the frame's, orientation = the base's — the angular rows
    are the relative angular velocity in BASE axes, which is what systemDynamicsVSMPC.cpp:170-183 multiplies by
    skew(wRb' axis);
  * centroidal momentum: about the CoM, world axes; Robot::getMomentum(true) rotates both halves by wRb' (:325-328);
  * mass matrix base block (6 x 6, mixed): [[m I, -m S(c)], [m S(c), I_o]], c = p_com - p_base, I_o the total inertia about
    the base origin in world axes.
"""
from __future__ import annotations

import numpy as np


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def rot_axis(axis, q):
    """Rodrigues: rotation by q about the unit axis."""
    K = skew(axis)
    return np.eye(3) + np.sin(q) * K + (1.0 - np.cos(q)) * (K @ K)


def as_rpy(R):
    """iDynTree::Rotation::asRPY."""
    if R[2, 0] < 1.0:
        if R[2, 0] > -1.0:
            return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arcsin(-R[2, 0]), np.arctan2(R[1, 0], R[0, 0])])
        return np.array([0.0, np.pi / 2.0, -np.arctan2(-R[1, 2], R[1, 1])])
    return np.array([0.0, -np.pi / 2.0, np.arctan2(-R[1, 2], R[1, 1])])


def forward_kinematics(model: dict, wRb, base_pos, q):
    """World rotation / origin of every link frame.  Link l >= 1 hangs on its parent through one revolute joint:
    parent_H_link(q) = [R0[l] Rot(axis[l], q[dof[l]]), p0[l]]."""
    n = model["n_links"]
    wR = [None] * n
    wp = [None] * n
    wR[0], wp[0] = np.asarray(wRb, float), np.asarray(base_pos, float)
    for l in range(1, n):
        p = model["parent"][l]
        wR[l] = wR[p] @ model["R0"][l] @ rot_axis(model["axis"][l], q[model["dof"][l]])
        wp[l] = wp[p] + wR[p] @ model["p0"][l]
    return wR, wp


def path_to_base(model, l):
    out = []
    while l > 0:
        out.append(l)
        l = model["parent"][l]
    return out            # links whose joint moves link l


def robot_set_state(model: dict, wRb, base_pos, base_lin_vel, omega_world, q, qd):
    """The kinematic rows of one instance's pack, as a dict of arrays named like paper_..._b200/pack.py (Jacobians with the
    controlled-joint columns only)."""
    n = model["n_links"]
    sel = list(model["sel"])
    wRb = np.asarray(wRb, float)
    wR, wp = forward_kinematics(model, wRb, base_pos, q)
    m = np.asarray(model["mass"], float)
    cw = [wp[l] + wR[l] @ model["com"][l] for l in range(n)]
    M = float(m.sum())
    p_com = sum(m[l] * cw[l] for l in range(n)) / M
    # per-link free-floating Jacobians of the link CoM (linear) and of the link (angular), joint columns (all dofs)
    nd = model["n_dof"]
    link_of_dof = {model["dof"][l]: l for l in range(1, n)}
    aw = {l: wR[l] @ model["axis"][l] for l in range(1, n)}
    Jc = np.zeros((n, 3, nd))
    Jw = np.zeros((n, 3, nd))
    for l in range(n):
        for k in path_to_base(model, l):
            j = model["dof"][k]
            Jc[l][:, j] = np.cross(aw[k], cw[l] - wp[k])
            Jw[l][:, j] = aw[k]
    J_com_full = sum(m[l] * Jc[l] for l in range(n)) / M
    # link velocities: base twist + joint part
    v0, w0 = np.asarray(base_lin_vel, float), np.asarray(omega_world, float)
    qd = np.asarray(qd, float)
    vc = [v0 + np.cross(w0, cw[l] - wp[0]) + Jc[l] @ qd for l in range(n)]
    wl = [w0 + Jw[l] @ qd for l in range(n)]
    Iw = [wR[l] @ model["inertia"][l] @ wR[l].T for l in range(n)]
    h_lin = sum(m[l] * vc[l] for l in range(n))
    h_ang = sum(Iw[l] @ wl[l] + m[l] * np.cross(cw[l] - p_com, vc[l]) for l in range(n))
    # mass matrix, base block (mixed): kinetic energy of the base twist alone
    c = p_com - wp[0]
    Io = sum(Iw[l] + m[l] * (skew(cw[l] - wp[0]).T @ skew(cw[l] - wp[0])) for l in range(n))
    M_b = np.zeros((6, 6))
    M_b[0:3, 0:3] = M * np.eye(3)
    M_b[0:3, 3:6] = -M * skew(c)
    M_b[3:6, 0:3] = M * skew(c)
    M_b[3:6, 3:6] = Io
    # jets
    p_off = p_com + wRb @ np.asarray(model["delta_com"], float)          # Robot.cpp:253-255
    jet_axes = np.zeros((4, 3))
    jet_arms = np.zeros((4, 3))
    J_jet_lin = np.zeros((4, 3, nd))
    J_rel_ang = np.zeros((4, 3, nd))
    A = np.zeros((6, 4))
    for i in range(4):
        l = model["jet_link"][i]
        pj = wp[l] + wR[l] @ model["jet_pos"][i]
        jet_axes[i] = wR[l] @ model["jet_axis"][i]
        jet_arms[i] = pj - p_off
        for k in path_to_base(model, l):
            j = model["dof"][k]
            J_jet_lin[i][:, j] = np.cross(aw[k], pj - wp[k])
            J_rel_ang[i][:, j] = wRb.T @ aw[k]
        A[0:3, i] = jet_axes[i]
        A[3:6, i] = np.cross(jet_arms[i], jet_axes[i])
    A_body = np.vstack([wRb.T @ A[0:3], wRb.T @ A[3:6]])
    return dict(
        wRb=wRb, omega_world=w0, rpy=as_rpy(wRb), mass=float(np.float32(M)), gravity=np.asarray(model["gravity"], float),
        M_b=M_b, base_pos=wp[0], p_com=p_com, momentum_body=np.concatenate([wRb.T @ h_lin, wRb.T @ h_ang]),
        A_mom_body=A_body, jet_axes=jet_axes, jet_arms=jet_arms, J_rel_ang=J_rel_ang[:, :, sel], J_jet_lin=J_jet_lin[:, :, sel],
        J_com=J_com_full[:, sel], J_com_full=J_com_full, J_jet_lin_full=J_jet_lin, J_rel_ang_full=J_rel_ang,
        link_rot=wR, link_pos=wp, h_lin=h_lin, h_ang=h_ang)
