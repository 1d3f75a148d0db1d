"""Drive the per-instance CPU oracle over a batch of synthetic states (test infrastructure)."""
import numpy as np

from helpers import fill_qp_input, load_trajectories, robot_data, set_robot_state
from oracle import vsmpc_oracle as O


class OracleInstance:
    """One reference-style MPC object + its QPInput/Robot, driven like src/variable_sampling_mpc.py."""

    def __init__(self, nominal_state, i, params=None, trajectories=None, qp_solver=None, jet_model=None, phase0=0):
        self.i = i
        self.params = dict(O.default_params())
        self.params.update(params or {})
        self.robot = robot_data(nominal_state, i)
        self.qp = O.QPInput()
        self.qp.setRobot(self.robot)
        self.qp.setRobotReference(self.robot)
        if jet_model is not None:
            self.qp.setJetModel(jet_model)
        else:
            self.qp.setEmptyJetModel()
        fill_qp_input(self.qp, nominal_state, i)
        self.mpc = O.VariableSamplingMPC(qp_solver=qp_solver)
        self.mpc.configure(self.params, self.qp, trajectories or load_trajectories(), phase0=phase0)

    def update(self, state):
        set_robot_state(self.robot, state, self.i)
        fill_qp_input(self.qp, state, self.i)
        self.mpc.update(self.qp)

    def solve(self):
        self.mpc.solveMPC()
        return self.mpc.getSolution()

    # dense pieces for K1 parity
    def dynamics(self):
        cs = self.mpc.vectorConstraints[0]
        return cs.A, cs.BJ, cs.BT, cs.c, cs.dt

    def output_row(self):
        m = self.mpc
        sel = m.jointSelectorVector
        return np.concatenate([m.deltaJointsPositionReference, m.getThrottleReference(), m.getThrustReference(),
                               m.getThrustDotReference(), m.finalState, m.getJointsReferencePosition()[sel]])


def oracle_trajectories_to_product(traj):
    """helpers.load_trajectories() layout -> product config layout."""
    a, p = traj["TRAJECTORY_MANAGER"], traj["POSITION_TRAJECTORY"]
    return dict(alpha_fps=a["fps"], alphaGravity=a["arrays"]["alphaGravity"], traj_fps=p["fps"],
                positionCoM=p["arrays"]["positionCoM"], velocityCoM=p["arrays"]["velocityCoM"],
                RPY=p["arrays"]["RPY"], RPYDot=p["arrays"]["RPYDot"])
