"""Test helpers: bridge the product's batched synthetic states to the per-instance oracle."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200"


def pkg(sub: str = ""):
    return importlib.import_module(PKG_NAME + (("." + sub) if sub else ""))


def load_trajectories():
    """Trajectory fixtures converted from the reference's .mat files (tests/golden/make_fixtures.py)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    return {
        "TRAJECTORY_MANAGER": dict(fps=int(d["alpha_fps"]), arrays={"alphaGravity": d["alphaGravity"]}),
        "POSITION_TRAJECTORY": dict(fps=int(d["traj_fps"]), arrays={
            "positionCoM": d["positionCoM"], "velocityCoM": d["velocityCoM"],
            "RPY": d["RPY"], "RPYDot": d["RPYDot"]}),
    }


def robot_data(state: dict, i: int):
    """Instance ``i`` of a getter-level batch -> oracle RobotData."""
    from oracle.vsmpc_oracle import RobotData
    return RobotData(
        wRb=state["wRb"][i].copy(), base_pos=state["base_pos"][i].copy(),
        omega_world=state["omega_world"][i].copy(), rpy=state["rpy"][i].copy(),
        mass_matrix_base=state["M_b"][i].copy(), p_com=state["p_com"][i].copy(),
        momentum_body=state["momentum_body"][i].copy(), A_mom_body=state["A_mom_body"][i].copy(),
        jet_axes=state["jet_axes"][i].copy(), jet_arms=state["jet_arms"][i].copy(),
        J_rel_body=state["J_rel_body"][i].copy(), J_jet_lin=state["J_jet_lin"][i].copy(),
        J_com=state["J_com"][i].copy(), jet_thrusts=state["thrust"][i].copy(),
        joint_pos=state["joint_pos"][i].copy(), gravity=state["gravity"][i].copy())


def set_robot_state(robot, state: dict, i: int):
    """In-place update of a RobotData (the reference mutates one shared Robot via setState)."""
    new = robot_data(state, i)
    for k, v in new.__dict__.items():
        setattr(robot, k, v)


def fill_qp_input(qp, state: dict, i: int):
    qp.setThrottleMPC(state["throttle_prev"][i])
    qp.setThrustDesMPC(state["thrust_des"][i])
    qp.setThrustDotDesMPC(state["thrust_dot_des"][i])
    qp.setEstimatedThrustDot(state["thrust_dot_est"][i])
    qp.setOutputQPJointsPosition(state["q_cmd"][i])


def state_from_pack(pack, joint_pos_sel=None, nJ=23, sel=tuple(range(3, 11))):
    """Rebuild a getter-level batch dict from a stored SoA pack (golden inputs).  Jacobian columns of
    the non-controlled joints are zero; they never enter the MPC (only the controlled columns are read)."""
    P = pkg("pack")
    B = pack.shape[1]
    f = lambda name, shape: np.ascontiguousarray(pack[P.PACK_OFFSETS[name][0]:sum(P.PACK_OFFSETS[name])].T).reshape((B,) + shape)
    J_rel_body = np.zeros((B, 4, 6, nJ))
    J_rel_body[:, :, 3:6, :][..., list(sel)] = f("J_rel_ang", (4, 3, 8))
    J_jet_lin = np.zeros((B, 4, 3, nJ))
    J_jet_lin[..., list(sel)] = f("J_jet_lin", (4, 3, 8))
    J_com = np.zeros((B, 3, nJ))
    J_com[..., list(sel)] = f("J_com", (3, 8))
    q_cmd = np.zeros((B, nJ))
    q_cmd[:, list(sel)] = f("q_cmd", (8,))
    joint_pos = np.zeros((B, nJ))
    if joint_pos_sel is not None:
        joint_pos[:, list(sel)] = np.asarray(joint_pos_sel).T
    return dict(
        wRb=f("wRb", (3, 3)), omega_world=f("omega_world", (3,)), rpy=f("rpy", (3,)), mass=f("mass", (1,))[:, 0],
        gravity=f("gravity", (3,)), M_b=f("M_b", (6, 6)), base_pos=f("base_pos", (3,)), p_com=f("p_com", (3,)),
        momentum_body=f("momentum_body", (6,)), A_mom_body=f("A_mom_body", (6, 4)), jet_axes=f("jet_axes", (4, 3)),
        jet_arms=f("jet_arms", (4, 3)), J_rel_body=J_rel_body, J_jet_lin=J_jet_lin, J_com=J_com,
        thrust=f("thrust", (4,)), thrust_dot_est=f("thrust_dot_est", (4,)), thrust_des=f("thrust_des", (4,)),
        thrust_dot_des=f("thrust_dot_des", (4,)), throttle_prev=f("throttle_prev", (4,)), q_cmd=q_cmd,
        joint_pos=joint_pos)


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))
