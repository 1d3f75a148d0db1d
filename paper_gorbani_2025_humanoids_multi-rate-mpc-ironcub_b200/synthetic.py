"""SYNTHETIC robot-state generator for tests and benchmarks.

None of the numbers below come from the reference: the iRonCub URDF (``ironcub-models 0.0.2``) is
not vendored in the reference repo, so the kinematic quantities the MPC reads from ``Robot``
(mass matrix, jet frames, relative Jacobians) are replaced by a plausible, seeded, *synthetic* robot
with the sparsity the real one has (``robot.toml``: base frame ``chest``; jets 0/1 on the arms,
jets 2/3 on the chest ⇒ chest-jet relative Jacobians are zero for the arm joints; each arm jet
depends only on its own arm's 4 joints) — SURVEY.md App. B-1.

The perturbation model is BASELINE/SURVEY §8(d) "Config 2".
"""
from __future__ import annotations

import numpy as np

NJ = 23
SEL = list(range(3, 11))


def rpy_to_R(rpy):
    """R = Rz(yaw) Ry(pitch) Rx(roll) — iDynTree::Rotation::RPY convention; rpy (..., 3)."""
    r, p, y = rpy[..., 0], rpy[..., 1], rpy[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty(rpy.shape[:-1] + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]
    S[..., 2, 1] = v[..., 0]
    return S


class SyntheticRobot:
    """Fixed (seeded) body-frame geometry of the synthetic robot."""

    def __init__(self, seed: int = 7):
        g = np.random.default_rng(seed)
        self.mass = 70.0
        self.I_body = np.diag([7.0, 6.0, 1.5])
        self.com_from_base_body = np.array([0.01, 0.0, -0.12])
        # jet positions relative to the CoM and thrust axes, body frame
        self.jet_pos_body = np.array([[0.02, 0.36, 0.18], [0.02, -0.36, 0.18],
                                      [-0.16, 0.11, 0.24], [-0.16, -0.11, 0.24]])
        ax = np.array([[0.05, 0.10, 1.0], [0.05, -0.10, 1.0], [-0.08, 0.02, 1.0], [-0.08, -0.02, 1.0]])
        self.jet_axes_body = ax / np.linalg.norm(ax, axis=1, keepdims=True)
        # relative Jacobians (body frame), joint columns; arm jet i depends on its own arm only
        self.J_rel_body = np.zeros((4, 6, NJ))
        self.J_rel_body[0][:, 3:7] = g.uniform(-0.35, 0.35, (6, 4))
        self.J_rel_body[1][:, 7:11] = g.uniform(-0.35, 0.35, (6, 4))
        self.J_rel_body[0][3:6, 3:7] = g.uniform(-1.0, 1.0, (3, 4))
        self.J_rel_body[1][3:6, 7:11] = g.uniform(-1.0, 1.0, (3, 4))
        # body-frame linear Jacobians of the jet frames and of the CoM (joint part)
        self.J_jet_lin_body = np.zeros((4, 3, NJ))
        self.J_jet_lin_body[0][:, 3:7] = self.J_rel_body[0][0:3, 3:7]
        self.J_jet_lin_body[1][:, 7:11] = self.J_rel_body[1][0:3, 7:11]
        self.J_com_body = g.uniform(-0.02, 0.02, (3, NJ))
        self.joint_pos0 = np.deg2rad(np.array([-0.0, -1.02, 0.0, -19.0, 18.68, 31.0, 15.0, -19.0, 18.68,
                                               31.0, 15.0, 19.6, 8.43, 4.64, 0.0, 1.71, -8.33, 19.6,
                                               8.43, 4.64, 0.0, 1.71, -8.33]))
        self.gravity = np.array([0.0, 0.0, -9.81])


def make_states(B: int, seed: int = 20251002, perturbed: bool = True, robot: SyntheticRobot = None,
                near_bound_fraction: float = 0.10, mass_scale=None, inertia_scale=None) -> dict:
    """Getter-level batch (arrays with leading dim B) — the data ``Robot``/``QPInput`` would return.

    ``perturbed=False`` gives the nominal hover state used for ``configure`` (so that the tracked
    reference is "hover at the perturbation-free initial CoM").
    """
    rb = robot or SyntheticRobot()
    g = np.random.default_rng(seed)
    z = lambda *s: np.zeros((B,) + s)
    p_com0 = np.array([0.0, 0.0, 1.0])
    if perturbed:
        p_com = p_com0 + g.normal(0, 0.05, (B, 3))
        lin_mom = g.normal(0, 2.0, (B, 3))
        rpy = g.normal(0, 0.05, (B, 3))
        ang_mom = g.normal(0, 0.5, (B, 3))
        thrust = g.uniform(60.0, 200.0, (B, 4))
        thrust_dot = g.normal(0, 20.0, (B, 4))
        throttle_prev = g.uniform(20.0, 90.0, (B, 4))
        omega_B = g.normal(0, 0.1, (B, 3))
        nb = int(round(near_bound_fraction * B))
        if nb > 0:
            idx = g.choice(B, nb, replace=False)
            lo = g.random((nb, 4)) < 0.5
            throttle_prev[idx] = np.where(lo, g.uniform(0.0, 1.0, (nb, 4)), g.uniform(99.0, 100.0, (nb, 4)))
        thrust_des = thrust + g.normal(0, 3.0, (B, 4))
        thrust_dot_des = g.normal(0, 10.0, (B, 4))
        dq = g.normal(0, 0.02, (B, NJ))
    else:
        p_com = np.tile(p_com0, (B, 1))
        lin_mom, rpy, ang_mom, omega_B = z(3), z(3), z(3), z(3)
        thrust = np.full((B, 4), rb.mass * 9.81 / 4.0 / 0.99)
        thrust_dot = z(4)
        throttle_prev = np.full((B, 4), 60.0)
        thrust_des, thrust_dot_des = thrust.copy(), z(4)
        dq = z(NJ)
    R = rpy_to_R(rpy)
    ms = np.ones(B) if mass_scale is None else np.asarray(mass_scale, float)
    isc = np.ones(B) if inertia_scale is None else np.asarray(inertia_scale, float)
    mass = np.float32(rb.mass * ms).astype(np.float64)  # Robot::m_totalMass is a float
    I_body = rb.I_body[None] * isc[:, None, None]
    c = np.einsum("bij,j->bi", R, rb.com_from_base_body)  # p_com - p_base, world
    Sc = skew(c)
    I_world = R @ I_body @ np.swapaxes(R, 1, 2)
    M_b = np.zeros((B, 6, 6))
    M_b[:, 0:3, 0:3] = mass[:, None, None] * np.eye(3)
    M_b[:, 0:3, 3:6] = -mass[:, None, None] * Sc
    M_b[:, 3:6, 0:3] = mass[:, None, None] * Sc
    M_b[:, 3:6, 3:6] = I_world + mass[:, None, None] * (np.swapaxes(Sc, 1, 2) @ Sc)
    jet_axes = np.einsum("bij,kj->bki", R, rb.jet_axes_body)
    jet_arms = np.einsum("bij,kj->bki", R, rb.jet_pos_body)
    A_mom = np.zeros((B, 6, 4))
    A_mom[:, 0:3, :] = np.swapaxes(jet_axes, 1, 2)
    A_mom[:, 3:6, :] = np.swapaxes(np.cross(jet_arms, jet_axes), 1, 2)
    Rt = np.swapaxes(R, 1, 2)
    A_mom_body = np.concatenate([Rt @ A_mom[:, 0:3, :], Rt @ A_mom[:, 3:6, :]], axis=1)
    J_jet_lin = np.einsum("bij,kjn->bkin", R, rb.J_jet_lin_body)
    J_com = np.einsum("bij,jn->bin", R, rb.J_com_body)
    state = dict(
        wRb=R, base_pos=p_com - c, omega_world=np.einsum("bij,bj->bi", R, omega_B), rpy=rpy,
        mass=mass, gravity=np.tile(rb.gravity, (B, 1)), M_b=M_b, p_com=p_com,
        momentum_body=np.concatenate([lin_mom, ang_mom], axis=1), A_mom_body=A_mom_body,
        jet_axes=jet_axes, jet_arms=jet_arms,
        J_rel_body=np.tile(rb.J_rel_body, (B, 1, 1, 1)), J_jet_lin=J_jet_lin, J_com=J_com,
        thrust=thrust, thrust_dot_est=thrust_dot, thrust_des=thrust_des, thrust_dot_des=thrust_dot_des,
        throttle_prev=throttle_prev, q_cmd=rb.joint_pos0[None, :] + dq,
        joint_pos=np.tile(rb.joint_pos0, (B, 1)),
    )
    return state


def apply_feedback(state: dict, out_rows: np.ndarray, sel=SEL) -> dict:
    """What the reference driver feeds back into QPInput after every tick
    (src/variable_sampling_mpc.py:124-131): throttle, desired thrust / thrust rate, joint references."""
    s = dict(state)
    s["throttle_prev"] = out_rows[:, 8:12].copy()
    s["thrust_des"] = out_rows[:, 12:16].copy()
    s["thrust_dot_des"] = out_rows[:, 16:20].copy()
    q = state["q_cmd"].copy()
    q[:, sel] = out_rows[:, 46:54]
    s["q_cmd"] = q
    return s
