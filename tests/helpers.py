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
    """Trajectory fixtures converted from the reference's .mat files (tools/make_fixtures.py)."""
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
