"""Single-instance mirror of the reference's Python bindings, running on the GPU through the C-ABI.

``VariableSamplingMPC`` and ``QPInput`` keep the method names, argument meaning and return values of
``momentum_based_mpc.bindingsMPC.VariableSamplingMPC`` (MPC/bindings/python/MPCPyBindings.cpp:22-90)
and of the ``QPInput`` setters the driver uses (UT/bindings/python/flightCtrlPyBindings.cpp;
src/variable_sampling_mpc.py:49-58,108-131), so that the reference's closed-loop script can swap the
import and nothing else:

    qpInput = QPInput(); qpInput.setRobot(robot); ...            # robot: RobotState (getter-level data)
    mpc = VariableSamplingMPC()
    mpc.configure(param_handler_mpc, qpInput)                    # dict from config.read_xml_config(...)
    mpc.update(qpInput); mpc.solveMPC()
    mpc.getThrustReference(); mpc.getThrottleReference(); mpc.getJointsReferencePosition()

``RobotState`` carries the outputs of the ``Robot`` getters the MPC reads; in the
reference they come from iDynTree (UT/src/Robot.cpp:198-335), which is outside this path.
"""
from __future__ import annotations

import os

import numpy as np

from .batched import BatchedVSMPC, VsmpcError
from .config import default_params, load_trajectories_mat, load_trajectories_npz
from .pack import DEFAULT_JOINT_SELECTOR

AXES_LIST = [  # src/config/robot.toml:3-27
    "torso_pitch", "torso_roll", "torso_yaw",
    "l_shoulder_pitch", "l_shoulder_roll", "l_shoulder_yaw", "l_elbow",
    "r_shoulder_pitch", "r_shoulder_roll", "r_shoulder_yaw", "r_elbow",
    "l_hip_pitch", "l_hip_roll", "l_hip_yaw", "l_knee", "l_ankle_pitch", "l_ankle_roll",
    "r_hip_pitch", "r_hip_roll", "r_hip_yaw", "r_knee", "r_ankle_pitch", "r_ankle_roll",
]


class RobotState:
    """Getter-level robot data for ONE robot (what ``Robot::setState`` leaves behind).

    Fields (all numpy arrays): wRb (3,3), base_pos (3), omega_world (3), rpy (3), M_b (6,6),
    p_com (3), momentum_body (6), A_mom_body (6,4), jet_axes (4,3), jet_arms (4,3),
    J_rel_body (4,6,nJ), J_jet_lin (4,3,nJ), J_com (3,nJ), thrust (4), joint_pos (nJ), gravity (3).
    """

    FIELDS = ("wRb", "base_pos", "omega_world", "rpy", "M_b", "p_com", "momentum_body", "A_mom_body",
              "jet_axes", "jet_arms", "J_rel_body", "J_jet_lin", "J_com", "thrust", "joint_pos", "gravity")

    def __init__(self, joint_names=None, **fields):
        self.joint_names = list(joint_names or AXES_LIST)
        for k in self.FIELDS:
            setattr(self, k, None)
        self.setState(**fields)

    def setState(self, **fields):
        for k, v in fields.items():
            if k not in self.FIELDS and k != "mass":
                raise VsmpcError(f"RobotState: unknown field {k}")
            setattr(self, k, np.asarray(v, dtype=np.float64))
        return True

    def getNJoints(self):
        return len(self.joint_names)

    def getJointName(self, i):
        return self.joint_names[i]

    def getJointPos(self):
        return self.joint_pos

    def getTotalMass(self) -> float:
        # Robot::m_totalMass is a float (UT/include/Robot.h:338)
        return float(np.float32(self.M_b[0, 0]))


class QPInput:
    """The QPInput setters/getters on the path (UT/include/QPInput.h:91-124)."""

    def __init__(self):
        self._robot = None
        self._robotReference = None
        self._throttleMPC = np.zeros(4)
        self._thrustDesMPC = np.zeros(4)
        self._thrustDotDesMPC = np.zeros(4)
        self._estimatedThrustDot = np.zeros(4)
        self._outputQPJointsPosition = None
        # written by the MPC (published references)
        self._posCoMReference = np.zeros(3)
        self._RPYReference = np.zeros(3)
        self._momentumReference = np.zeros(6)
        self._alphaGravity = 0.0

    def setRobot(self, r): self._robot = r
    def setRobotReference(self, r): self._robotReference = r
    def getRobot(self): return self._robot
    def getRobotReference(self): return self._robotReference
    def setEmptyVectorsCollectionServer(self): pass
    def setEmptyJetModel(self): pass
    def setThrottleMPC(self, v): self._throttleMPC = np.array(v, dtype=np.float64).reshape(4)
    def getThrottleMPC(self): return self._throttleMPC
    def setThrustDesMPC(self, v): self._thrustDesMPC = np.array(v, dtype=np.float64).reshape(4)
    def getThrustDesMPC(self): return self._thrustDesMPC
    def setThrustDotDesMPC(self, v): self._thrustDotDesMPC = np.array(v, dtype=np.float64).reshape(4)
    def getThrustDotDesMPC(self): return self._thrustDotDesMPC
    def setEstimatedThrustDot(self, v): self._estimatedThrustDot = np.array(v, dtype=np.float64).reshape(4)
    def getEstimatedThrustDot(self): return self._estimatedThrustDot
    def setOutputQPJointsPosition(self, v): self._outputQPJointsPosition = np.array(v, dtype=np.float64)
    def getOutputQPJointsPosition(self): return self._outputQPJointsPosition
    def getPosCoMReference(self): return self._posCoMReference
    def getRPYReference(self): return self._RPYReference
    def getMomentumReference(self): return self._momentumReference
    def getAlphaGravity(self): return self._alphaGravity
    # written by the MPC's update (costsVSMPC.cpp:155-160, systemDynamicsVSMPC.cpp:310)
    def setPosCoMReference(self, v): self._posCoMReference = np.array(v, dtype=np.float64).reshape(3)
    def setRPYReference(self, v): self._RPYReference = np.array(v, dtype=np.float64).reshape(3)
    def setMomentumReference(self, v): self._momentumReference = np.array(v, dtype=np.float64).reshape(6)
    def setAlphaGravity(self, v): self._alphaGravity = float(v)


def _state_dict(qp: QPInput) -> dict:
    r = qp.getRobot()
    if r is None or qp.getRobotReference() is None:
        raise VsmpcError("QPInput: setRobot / setRobotReference first")
    if qp.getRobotReference() is not r:
        raise VsmpcError("robot and robotReference must be the same object (as in src/variable_sampling_mpc.py:50-51)")
    qcmd = qp.getOutputQPJointsPosition()
    if qcmd is None:
        raise VsmpcError("QPInput: setOutputQPJointsPosition first")
    one = lambda a: np.asarray(a, dtype=np.float64)[None]
    return dict(
        wRb=one(r.wRb), base_pos=one(r.base_pos), omega_world=one(r.omega_world), rpy=one(r.rpy),
        mass=np.array([r.getTotalMass()]), gravity=one(r.gravity), M_b=one(r.M_b), p_com=one(r.p_com),
        momentum_body=one(r.momentum_body), A_mom_body=one(r.A_mom_body), jet_axes=one(r.jet_axes),
        jet_arms=one(r.jet_arms), J_rel_body=one(r.J_rel_body), J_jet_lin=one(r.J_jet_lin), J_com=one(r.J_com),
        thrust=one(r.thrust), thrust_dot_est=one(qp.getEstimatedThrustDot()), thrust_des=one(qp.getThrustDesMPC()),
        thrust_dot_des=one(qp.getThrustDotDesMPC()), throttle_prev=one(qp.getThrottleMPC()), q_cmd=one(qcmd),
        joint_pos=one(r.joint_pos))


class VariableSamplingMPC:
    """Drop-in for ``momentum_based_mpc.bindingsMPC.VariableSamplingMPC`` (one instance, device 0)."""

    def __init__(self, device: int = 0):
        self._device = device
        self._impl = None
        self._robot = None
        self._sel = None
        self._jointsPositionReference = None

    @staticmethod
    def _params(handler) -> dict:
        if isinstance(handler, dict):
            return handler
        # a BLF-like parameters handler: get_parameter_* accessors
        p = {}
        for k, v in default_params().items():
            for getter in ("get_parameter_float", "get_parameter_int", "get_parameter_bool",
                           "get_parameter_vector_float", "get_parameter_vector_string", "get_parameter_string"):
                try:
                    p[k] = getattr(handler, getter)(k)
                    break
                except Exception:
                    continue
            else:
                raise VsmpcError(f"Parameter '{k}' not found in the config file.")
        return p

    def configure(self, parametersHandler, qpInput: QPInput, trajectories: dict | None = None) -> bool:
        p = dict(self._params(parametersHandler))
        robot = qpInput.getRobot()
        names = list(p.get("controlledJoints", default_params()["controlledJoints"]))
        if len(names) != 8:
            # variableSamplingMPC.cpp:18-23
            raise VsmpcError("The number of controlled joints defined in the systemDynamic.h file is different "
                             "from the size of the 'controlledJoints' parameter")
        self._sel = [j for n in names for j in range(robot.getNJoints()) if n == robot.getJointName(j)]
        if len(self._sel) != 8:
            raise VsmpcError("controlledJoints not found in the robot's joint list")
        if trajectories is None:
            tm, pt = p.get("TRAJECTORY_MANAGER", {}), p.get("POSITION_TRAJECTORY", {})
            a, b = tm.get("trajectoryFile"), pt.get("trajectoryFile")
            if not a or not b:
                raise VsmpcError("Group [TRAJECTORY_MANAGER] not found in the config file.")
            if a.endswith(".npz"):
                trajectories = load_trajectories_npz(a)
            else:
                for f in (a, b):
                    if not os.path.exists(f):
                        raise VsmpcError(f"Error opening file {f}")
                trajectories = load_trajectories_mat(a, b)
        self._impl = BatchedVSMPC(1, p, trajectories, device=self._device, full_solution=True)
        self._impl.sel = list(self._sel)
        self._robot = robot
        self._qp = qpInput
        self._jointsPositionReference = np.array(robot.getJointPos(), dtype=np.float64).copy()
        self._impl.configure(_state_dict(qpInput))
        self._publish(qpInput)
        self._status = 0
        self._out = np.zeros(54)
        self._out[46:54] = self._jointsPositionReference[self._sel]
        return True

    def _publish(self, qpInput: QPInput):
        """The QPInput fields update() writes: tracked references (costsVSMPC.cpp:155-160) and the gravity-compensation
        factor (systemDynamicsVSMPC.cpp:310), read back from the device-resident tick state."""
        r = self._impl.get_references()
        qpInput.setPosCoMReference(r["posCoMReference"][0])
        qpInput.setRPYReference(r["RPYReference"][0])
        qpInput.setMomentumReference(r["momentumReference"][0])
        qpInput.setAlphaGravity(r["alphaGravity"][0])

    def update(self, qpInput: QPInput) -> bool:
        self._impl.update(_state_dict(qpInput))
        self._publish(qpInput)
        return True

    def solveMPC(self) -> bool:
        self._impl.solveMPC()
        out, status = self._impl.get_output()
        self._out, self._status = out[0], int(status[0])
        # joint accumulator lives on the device for the controlled joints; the others never change
        self._jointsPositionReference[self._sel] = self._out[46:54]
        return True  # the reference always returns true (variableSamplingMPC.cpp:111)

    # ---- getters (MPCPyBindings.cpp:38-90) ------------------------------------------------------------
    def getJointsReferencePosition(self): return self._jointsPositionReference.copy()
    def getThrottleReference(self): return self._out[8:12].copy()
    def getThrustReference(self): return self._out[12:16].copy()
    def getThrustDotReference(self): return self._out[16:20].copy()
    def getFinalCoMPosition(self): return self._out[20:23].copy()
    def getFinalLinMom(self): return self._out[23:26].copy()
    def getFinalRPY(self): return self._out[26:29].copy()
    def getFinalAngMom(self): return self._out[29:32].copy()
    def getNStatesMPC(self): return 26.0
    def getNInputMPC(self): return 12.0
    def getSolution(self): return self._impl.getSolution()[0]
    def getQPProblemStatus(self): return self._status
    def getNOptimizationVariables(self): return self._impl.n_var
    def getNConstraints(self): return self._impl.n_con
    def getGradient(self): return self._impl.get_qp_vectors()[0][0]
    def getLowerBound(self): return self._impl.get_qp_vectors()[1][0]
    def getUpperBound(self): return self._impl.get_qp_vectors()[2][0]
    def getHessian(self): return self._impl.getHessian(0)
    def getLinearConstraintMatrix(self): return self._impl.getLinearConstraintMatrix(0)
