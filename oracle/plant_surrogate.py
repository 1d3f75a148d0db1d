"""TEST INFRASTRUCTURE — CPU restatement (NumPy) of the surrogate plant used for closed-loop parity.

The reference closes the loop through MuJoCo (src/mujoco_lib/ironcub_mujoco_simulator.py), which is not
available here; both the CUDA rollout (csrc/vsmpc_plant.cu) and this file integrate the SAME
stated surrogate: the MPC's own nonlinear model with frozen body-frame kinematics,
  jets      : jet_kalman_filter.py:30-45  (Td += sigma_T (f + g v(u)) dt ; T += Td dt)
  momentum  : h_lin^w' = alpha_g m g + sum_i (T_i + dT_i) R a_i ;  h_ang^B' = -w_B x h_ang^B + sum_i (T_i + dT_i) r_i x a_i
  pose      : p' = h_lin^w / m ;  rpy' = W^-1(rpy) I_B^-1 h_ang^B
  joints    : q = q_cmd (position control); the jet frames follow the controlled joints through FIRST-ORDER kinematics
              about the configure-time posture q0 with frozen relative Jacobians (the same sensitivities the MPC's
              Lambda matrices are built from, systemDynamicsVSMPC.cpp:159-226,321-350):
              a_i(q) = a_i0 + (J^w_rel,i dq) x a_i0 ,  r_i(q) = r_i0 + (J^lin_i - J^CoM) dq ,  dq = q - q0
``SurrogateLoop`` drives one oracle MPC instance in closed loop exactly like the reference driver
(src/variable_sampling_mpc.py:106-131).  Written independently of the product's Python (it shares no code
with paper_..._b200/synthetic.py or rollout.py); only tests/ and bench.py's cpu_baseline may import it.
"""
from __future__ import annotations

import math

import numpy as np

from .vsmpc_oracle import JetModel, RobotData


def _R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _S(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


class SurrogatePlant:
    def __init__(self, geometry: dict, mass: float, I_body, thrust_disturbance, state: dict, dt_sim=0.001, n_sub=5,
                 jet_model=None, jet_nn=None, ekf_R=None, ekf_Q=None):
        """geometry: com_from_base_body(3), jet_pos_body(4,3), jet_axes_body(4,3), J_rel_body(4,6,nJ),
        J_jet_lin_body(4,3,nJ), J_com_body(3,nJ), gravity(3), joint_pos0(nJ).  state: p_com, lin_mom_world, rpy,
        ang_mom_body, thrust, thrust_dot, throttle, thrust_des, thrust_dot_des, q_cmd(nJ)."""
        self.g = geometry
        self.mass = float(np.float32(mass))
        self.I = np.array(I_body, float)
        self.Iinv = np.linalg.inv(self.I)
        self.dT = np.array(thrust_disturbance, float)
        self.s = {k: np.array(v, float) for k, v in state.items()}
        self.dt, self.n_sub = dt_sim, n_sub
        self.jet = jet_model or JetModel()
        self.sel = list(range(3, 11))
        # jet-NN mode: neural jet plant + per-jet EKF of the reference simulator
        # (ironcub_mujoco_simulator.py:50-57,129-133)
        self.jet_nn = jet_nn
        if jet_nn is not None:
            from .jet_nn_oracle import JetEKF
            R = np.eye(2) * 0.5 if ekf_R is None else ekf_R
            Q = np.eye(2) * 0.1 if ekf_Q is None else ekf_Q
            self.ekf = [JetEKF(R, Q, np.eye(2) * 0.1, dt_sim, self.jet) for _ in range(4)]
            self.T_nn = np.asarray(self.s["thrust"], np.float32).copy()

    def jet_frames_body(self):
        """Thrust axes / arms in the body frame at the current joint command (first-order kinematics)."""
        g = self.g
        dq = (self.s["q_cmd"] - g["joint_pos0"])[self.sel]
        a = np.empty((4, 3)); r = np.empty((4, 3))
        for i in range(4):
            w = g["J_rel_body"][i][3:6][:, self.sel] @ dq
            a[i] = g["jet_axes_body"][i] + np.cross(w, g["jet_axes_body"][i])
            r[i] = g["jet_pos_body"][i] + (g["J_jet_lin_body"][i][:, self.sel] - g["J_com_body"][:, self.sel]) @ dq
        return a, r

    def step(self, alpha_g: float = 1.0):
        """alpha_g: the gravity-compensation factor the MPC published this tick (the ground carries 1 - alpha_g)."""
        s, g, jet = self.s, self.g, self.jet
        sig = jet.getThrustStandardDeviation_u2T()
        aB, rB = self.jet_frames_body()
        for _ in range(self.n_sub):
            R = _R(s["rpy"])
            if self.jet_nn is not None:
                from .jet_nn_oracle import nn_jet_step
                u32 = np.asarray(s["throttle"], np.float32)          # the simulator keeps the throttle in float32
                self.T_nn, Tdn = nn_jet_step(self.T_nn, u32, self.jet_nn, self.dt)
                for j in range(4):
                    x = self.ekf[j].update([s["thrust"][j], s["thrust_dot"][j]], float(u32[j]), [float(self.T_nn[j]), float(Tdn[j])])
                    s["thrust"][j], s["thrust_dot"][j] = x[0], x[1]
            else:
                for j in range(4):
                    Ts, Tds = jet.standardizeThrust_u2T(s["thrust"][j]), jet.standardizeThrustDot_u2T(s["thrust_dot"][j])
                    v = jet.compute_v(jet.standardizeThrottle_u2T(s["throttle"][j]))
                    tdd = jet.compute_f(Ts, Tds) + jet.compute_g(Ts, Tds) * v
                    s["thrust_dot"][j] += tdd * sig * self.dt
                    s["thrust"][j] += s["thrust_dot"][j] * self.dt
            Tt = s["thrust"] + self.dT
            fB = Tt @ aB
            tauB = Tt @ np.cross(rB, aB)
            wB = self.Iinv @ s["ang_mom_body"]
            s["lin_mom_world"] = s["lin_mom_world"] + self.dt * (alpha_g * self.mass * g["gravity"] + R @ fB)
            s["ang_mom_body"] = s["ang_mom_body"] + self.dt * (tauB - np.cross(wB, s["ang_mom_body"]))
            wB = self.Iinv @ s["ang_mom_body"]
            r0, r1 = s["rpy"][0], s["rpy"][1]
            s0, c0, t1, c1 = math.sin(r0), math.cos(r0), math.tan(r1), math.cos(r1)
            Wi = np.array([[1, s0 * t1, c0 * t1], [0, c0, -s0], [0, s0 / c1, c0 / c1]])
            s["p_com"] = s["p_com"] + self.dt * s["lin_mom_world"] / self.mass
            s["rpy"] = s["rpy"] + self.dt * (Wi @ wB)

    def robot_data(self) -> RobotData:
        """What ``Robot::setState`` would leave behind for this plant state."""
        s, g = self.s, self.g
        R = _R(s["rpy"])
        c = R @ g["com_from_base_body"]
        Sc = _S(c)
        M = np.zeros((6, 6))
        M[:3, :3] = self.mass * np.eye(3)
        M[:3, 3:] = -self.mass * Sc
        M[3:, :3] = self.mass * Sc
        M[3:, 3:] = R @ self.I @ R.T + self.mass * (Sc.T @ Sc)
        aB, rB = self.jet_frames_body()
        aw = aB @ R.T
        rw = rB @ R.T
        A = np.zeros((6, 4))
        A[:3] = (aw @ R).T
        A[3:] = (np.cross(rw, aw) @ R).T
        wB = self.Iinv @ s["ang_mom_body"]
        return RobotData(
            wRb=R, base_pos=s["p_com"] - c, omega_world=R @ wB, rpy=s["rpy"].copy(), mass_matrix_base=M,
            p_com=s["p_com"].copy(), momentum_body=np.concatenate([R.T @ s["lin_mom_world"], s["ang_mom_body"]]),
            A_mom_body=A, jet_axes=aw, jet_arms=rw, J_rel_body=g["J_rel_body"].copy(),
            J_jet_lin=np.einsum("ij,kjn->kin", R, g["J_jet_lin_body"]), J_com=R @ g["J_com_body"],
            jet_thrusts=s["thrust"].copy(), joint_pos=g["joint_pos0"].copy(), gravity=g["gravity"].copy())


class SurrogateLoop:
    """One closed loop: oracle MPC + surrogate plant, sequenced like src/variable_sampling_mpc.py:68-71,106-161."""

    def __init__(self, plant: SurrogatePlant, trajectories=None, params=None, sel=tuple(range(3, 11))):
        from . import vsmpc_oracle as O
        self.O = O
        self.plant = plant
        self.sel = list(sel)
        self.robot = plant.robot_data()
        self.qp = O.QPInput()
        self.qp.setRobot(self.robot)
        self.qp.setRobotReference(self.robot)
        self.qp.setJetModel(plant.jet)
        self._feed()
        self.mpc = O.VariableSamplingMPC()
        p = dict(O.default_params())
        p.update(params or {})
        assert self.mpc.configure(p, self.qp, trajectories)

    def _feed(self):
        s = self.plant.s
        self.qp.setThrottleMPC(s["throttle"])
        self.qp.setThrustDesMPC(s["thrust_des"])
        self.qp.setThrustDotDesMPC(s["thrust_dot_des"])
        self.qp.setEstimatedThrustDot(s["thrust_dot"])
        self.qp.setOutputQPJointsPosition(s["q_cmd"])

    def tick(self):
        new = self.plant.robot_data()                       # sim.update_robot_state()
        for k, v in new.__dict__.items():
            setattr(self.robot, k, v)
        self.qp.setEstimatedThrustDot(self.plant.s["thrust_dot"])
        self.mpc.update(self.qp)
        self.mpc.solveMPC()
        s = self.plant.s
        s["throttle"] = np.array(self.mpc.getThrottleReference(), float)
        s["thrust_des"] = np.array(self.mpc.getThrustReference(), float)
        s["thrust_dot_des"] = np.array(self.mpc.getThrustDotReference(), float)
        s["q_cmd"] = np.array(self.mpc.getJointsReferencePosition(), float)
        self._feed()
        self.plant.step(self.qp.getAlphaGravity())          # sim.step(n_steps)
        return np.concatenate([s["p_com"], s["rpy"], s["thrust"], s["throttle"]])
