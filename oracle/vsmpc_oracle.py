"""CPU oracle (NumPy, FP64) for the variable-sampling MPC hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, class by class, what the reference computes on the path
``VariableSamplingMPC::update()`` + ``solveMPC()`` so that the CUDA path can be checked against it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it; the
product path (``paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200``) never does.

PARITY: PINNED against the reference's own code for everything except the OSQP iteration.  The reference
ships no tests, golden vectors or logs for this path and its third-party stack (Eigen,
OSQP 1.0.0 / QDLDL 0.1.8 via osqp-eigen 0.11.0, iDynTree 14.0.2, BLF, YARP, matio) is not in the build
container — but its 13 hot-path translation units compile, where they lie, against the stand-in headers of
oracle/ref_stubs/ (oracle/build_ref.py -> oracle/_ref/libvsmpc_reference.so), and tests/test_reference_pinned.py
holds this file to what that library returns (dense P, q, A, l, u at 1e-12; minimiser and outputs at 1e-9;
frozen in tests/golden/reference_{qp,ticks}.npz).  Further pins: the compiled ``UT/src/JetModel.cpp``
(tests/golden/jet_model_ref.npz), the reference's fixtures (``src/trajectories/*.mat``), the second
statement of the jet model in ``src/mujoco_lib/jet_kalman_filter.py:6-45``, closed-form identities
(tests/test_oracle.py).  UNPINNED: the OSQP iteration (restated in oracle/c/vsmpc_ref.c from the published
algorithm; the exact solver below is the arbiter of parity), iDynTree's kinematics (data here).

Reference paths (``MPC/`` = src/flight-controller/momentum-based-linear-mpc-lib,
``UT/`` = src/flight-controller/utils):

* JetModel                    UT/src/JetModel.cpp:10-114
* TrajectoryManager           UT/src/TrajectoryManager.cpp:23-167
* ReferenceTrackingCost etc.  MPC/src/variableSamplingMPC/costsVSMPC.cpp
* SystemDynamicVS (+3 parts)  MPC/src/variableSamplingMPC/systemDynamicsVSMPC.cpp
* constraints                 MPC/src/variableSamplingMPC/constraintsVSMPC.cpp, MPC/src/IMPCProblem/IQPUtilsMPC.cpp:57-92
* IMPCProblem                 MPC/src/IMPCProblem/IMPCProblem.cpp:3-298
* VariableSamplingMPC         MPC/src/variableSamplingMPC/variableSamplingMPC.cpp:7-227

Rigid-body quantities the reference obtains from iDynTree through ``Robot`` (UT/src/Robot.cpp:198-335)
are *data* here (class ``RobotData``), exactly the getters the path calls.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# --- MPC/include/variableSamplingMPC/VSconstant.h:6-16 ------------------------------------------
N_JOINTS = 8
N_THRUSTS = 4
CoMPosIdx = (0, 1, 2)
linMomIdx = (3, 4, 5)
rpyIdx = (6, 7, 8)
angMomIdx = (9, 10, 11)
thrustIdx = (12, 13, 14, 15)
thrustDotIdx = (16, 17, 18, 19)
positionErrorIdx = (20, 21, 22)
rpyErrorIdx = (23, 24, 25)

AXES_LIST = [  # src/config/robot.toml:3-27
    "torso_pitch", "torso_roll", "torso_yaw",
    "l_shoulder_pitch", "l_shoulder_roll", "l_shoulder_yaw", "l_elbow",
    "r_shoulder_pitch", "r_shoulder_roll", "r_shoulder_yaw", "r_elbow",
    "l_hip_pitch", "l_hip_roll", "l_hip_yaw", "l_knee", "l_ankle_pitch", "l_ankle_roll",
    "r_hip_pitch", "r_hip_roll", "r_hip_yaw", "r_knee", "r_ankle_pitch", "r_ankle_roll",
]
JETS_LIST = ["l_arm_jet_turbine", "r_arm_jet_turbine", "chest_l_jet_turbine", "chest_r_jet_turbine"]


def default_params() -> dict:
    """src/config/vs_mcp_config.xml:5-44 (group VS_MPC_CONFIG)."""
    return dict(
        useJetDynamic=True,
        useEstimatedThrust=True,
        periodMPC=0.005,
        periodMPCLargeSteps=0.1,
        periodMPCSmallSteps=0.005,
        nIter=17,
        nIterSmall=7,
        controlHorizon=12,
        controlledJoints=["l_shoulder_pitch", "l_shoulder_roll", "l_shoulder_yaw", "l_elbow",
                          "r_shoulder_pitch", "r_shoulder_roll", "r_shoulder_yaw", "r_elbow"],
        jointsLambdaOption="unfiltered",
        weightCoMPos=[500.0, 500.0, 5000.0],
        weightCoMPosError=[25000.0, 25000.0, 50000.0],
        weightLinMom=[1.0, 1.0, 1.5],
        weightRPY=[1000.0, 1000.0, 1000.0],
        weightRPYError=[10000.0, 10000.0, 10000.0],
        weightAngMom=[80.0, 80.0, 80.0],
        weightDeltaJoint=[65000.0] * 8,
        weightThrottle=80000.0,
        weightInitialThrottle=80000.0,
        weightRegularizationJointPos=20.0,
        throttleMin=0.0,
        throttleMax=100.0,
    )


def from_vec_to_skew(v: np.ndarray) -> np.ndarray:
    """UT/src/FlightControlUtils.cpp:77-85."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def adjoint_transform(R: np.ndarray, p: np.ndarray) -> np.ndarray:
    """iDynTree::Transform::asAdjointTransform for H=(R,p): [[R, S(p)R],[0, R]] (used at
    systemDynamicsVSMPC.cpp:123-130, costsVSMPC.cpp:283-285)."""
    X = np.zeros((6, 6))
    X[0:3, 0:3] = R
    X[0:3, 3:6] = from_vec_to_skew(p) @ R
    X[3:6, 3:6] = R
    return X


# =================================================================================================
class JetModel:
    """UT/src/JetModel.cpp:10-114.  Coefficients may be overridden per instance (config 5)."""

    def __init__(self, coeff=None, normalization=None):
        self.c = np.array([-4.64730485e-01, -8.13171858e+00, -6.19539230e+00, 6.61113140e-01,
                           1.67673231e+00, -4.83287064e-01, 8.77996617e+00, -1.01096376e+00,
                           -5.86442286e-01, 5.19093322e-01, -4.23782666e-01, -1.45705257e+00,
                           -7.83052261e-03]) if coeff is None else np.asarray(coeff, float)
        self.n = np.array([108.309, 65.793, 47.333, 31.483]) if normalization is None \
            else np.asarray(normalization, float)

    def compute_f(self, T, Td):
        c = self.c
        return c[0] + c[1] * T + c[2] * Td + c[3] * T * Td + c[4] * T ** 2.0 + c[5] * Td ** 2.0

    def compute_df_dT(self, T, Td):
        c = self.c
        return c[1] + c[3] * Td + 2 * c[4] * T

    def compute_df_dTdot(self, T, Td):
        c = self.c
        return c[2] + c[3] * T + 2 * c[5] * Td

    def compute_dg_dT(self, T, Td):
        c = self.c
        return c[7] + c[9] * Td + 2 * c[10] * T

    def compute_dg_dTdot(self, T, Td):
        c = self.c
        return c[8] + c[9] * T + 2 * c[11] * Td

    def compute_g(self, T, Td):
        c = self.c
        return c[6] + c[7] * T + c[8] * Td + c[9] * T * Td + c[10] * T ** 2.0 + c[11] * Td ** 2.0

    def compute_v(self, u):
        return u + self.c[12] * u ** 2.0

    def standardizeThrust_u2T(self, thrust):
        return (thrust - self.n[0]) / self.n[1]

    def standardizeThrustDot_u2T(self, thrustDot):
        return thrustDot / self.n[1]

    def standardizeThrottle_u2T(self, throttle):
        return (throttle - self.n[2]) / self.n[3]

    def destandardizeThrottle_u2T(self, v):
        d = 1 + 4 * self.c[12] * v
        # std::sqrt of a negative argument is NaN and both comparisons below are then false (JetModel.cpp:97-108);
        # unreachable from the MPC, whose v is boxed to [v(0 %), v(100 %)]
        u = (-1 + (math.sqrt(d) if d >= 0 else math.nan)) / (2 * self.c[12])
        u = u * self.n[3] + self.n[2]
        if u < 0:
            u = 0.0
        elif u > 100:
            u = 100.0
        return u

    def getThrustStandardDeviation_u2T(self):
        return self.n[1]


# =================================================================================================
class TrajectoryManager:
    """Array-backed stand-in for UT/src/TrajectoryManager.cpp (matio loader replaced by arrays).

    ``arrays`` maps key -> (dim, n_samples); ``fps`` is the file's fps; ``des_fps`` follows the
    reference's ``int`` parameter (float→int truncation at the call sites,
    systemDynamicsVSMPC.cpp:272, costsVSMPC.cpp:68).
    """

    def __init__(self, arrays: Dict[str, np.ndarray], fps: int, des_fps: float):
        des_fps = int(des_fps)  # implicit double→int conversion of the C++ call
        self.trajectories = {}
        self.trajectorySize = 0
        self.trajectoryIndex = 0
        for name, arr in arrays.items():
            vals = np.asarray(arr, float)
            if fps != des_fps and vals.shape[1] > 1:
                vals = self._upsample(vals, fps, des_fps)
            self.trajectories[name] = vals
            self.trajectorySize = max(self.trajectorySize, vals.shape[1])

    @staticmethod
    def _upsample(values: np.ndarray, fps: int, des_fps: int) -> np.ndarray:
        """Trajectory::upsample, TrajectoryManager.cpp:23-39 (drops the final sample)."""
        ratio = float(des_fps) / fps
        out = []
        for i in range(values.shape[1] - 1):
            k = 0
            while k < ratio:
                out.append(values[:, i] + (values[:, i + 1] - values[:, i]) * (k / ratio))
                k += 1
        return np.array(out).T

    def advanceTrajectory(self):
        if self.trajectoryIndex < self.trajectorySize - 1:
            self.trajectoryIndex += 1
        return True

    def getCurrentValue(self, key: str) -> np.ndarray:
        return self.trajectories[key][:, self.trajectoryIndex]


# =================================================================================================
@dataclass
class RobotData:
    """What the path reads from ``Robot`` (UT/include/Robot.h) after ``Robot::setState``.

    All members are the *outputs* of iDynTree-backed getters, supplied as data.
    """
    wRb: np.ndarray                 # getBasePose().getRotation()              (3,3)
    base_pos: np.ndarray            # getBasePose().getPosition()              (3,)
    omega_world: np.ndarray         # getBaseVel().getAngularVec3()            (3,)
    rpy: np.ndarray                 # getBasePose().getRotation().asRPY()      (3,)
    mass_matrix_base: np.ndarray    # getMassMatrix().block(0,0,6,6)           (6,6)
    p_com: np.ndarray               # getPositionCoM()                         (3,)
    momentum_body: np.ndarray       # getMomentum(true)                        (6,)
    A_mom_body: np.ndarray          # getMatrixAmomJets(true)                  (6,4)
    jet_axes: np.ndarray            # getMatrixOfJetAxes()[i]  (world)         (4,3)
    jet_arms: np.ndarray            # getMatrixOfJetArms()[i]  (world)         (4,3)
    J_rel_body: np.ndarray          # getRelativeJacobianJetsBodyFrame()[i]    (4,6,nJ)
    J_jet_lin: np.ndarray           # getJacobian(frame).topRightCorner(3,nJ)  (4,3,nJ)
    J_com: np.ndarray               # getJacobianCoM().topRightCorner(3,nJ)    (3,nJ)
    jet_thrusts: np.ndarray         # getJetThrusts()                          (4,)
    joint_pos: np.ndarray           # getJointPos()                            (nJ,)
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    joint_names: List[str] = field(default_factory=lambda: list(AXES_LIST))
    jets_list: List[str] = field(default_factory=lambda: list(JETS_LIST))

    def getNJoints(self):
        return len(self.joint_names)

    def getNJets(self):
        return len(self.jets_list)

    def getTotalMass(self) -> float:
        # Robot::m_totalMass is a float (UT/include/Robot.h:338; Robot.cpp:332)
        return float(np.float32(self.mass_matrix_base[0, 0]))


class QPInput:
    """The fields of UT/include/QPInput.h:91-124 the path touches."""

    def __init__(self):
        self.robot: Optional[RobotData] = None
        self.robotReference: Optional[RobotData] = None
        self.jetModel: Optional[JetModel] = None
        self.posCoMReference = np.zeros(3)
        self.RPYReference = np.zeros(3)
        self.momentumReference = np.zeros(6)
        self.alphaGravity = 0.0
        self.thrustDesMPC = np.zeros(4)
        self.thrustDotDesMPC = np.zeros(4)
        self.throttleMPC = np.zeros(4)
        self.estimatedThrustDot = np.zeros(4)
        self.outputQPJointsPosition = np.zeros(23)

    # setters/getters with the reference's names (flightCtrlPyBindings.cpp)
    def setRobot(self, r): self.robot = r
    def setRobotReference(self, r): self.robotReference = r
    def getRobot(self): return self.robot
    def getRobotReference(self): return self.robotReference
    def setEmptyJetModel(self): self.jetModel = JetModel()
    def setJetModel(self, m): self.jetModel = m
    def getJetModel(self): return self.jetModel
    def setEmptyVectorsCollectionServer(self): pass
    def setThrottleMPC(self, v): self.throttleMPC = np.array(v, float)
    def getThrottleMPC(self): return self.throttleMPC
    def setThrustDesMPC(self, v): self.thrustDesMPC = np.array(v, float)
    def getThrustDesMPC(self): return self.thrustDesMPC
    def setThrustDotDesMPC(self, v): self.thrustDotDesMPC = np.array(v, float)
    def getThrustDotDesMPC(self): return self.thrustDotDesMPC
    def setEstimatedThrustDot(self, v): self.estimatedThrustDot = np.array(v, float)
    def getEstimatedThrustDot(self): return self.estimatedThrustDot
    def setOutputQPJointsPosition(self, v): self.outputQPJointsPosition = np.array(v, float)
    def getOutputQPJointsPosition(self): return self.outputQPJointsPosition
    def setPosCoMReference(self, v): self.posCoMReference = np.array(v, float)
    def getPosCoMReference(self): return self.posCoMReference
    def setRPYReference(self, v): self.RPYReference = np.array(v, float)
    def getRPYReference(self): return self.RPYReference
    def setMomentumReference(self, v): self.momentumReference = np.array(v, float)
    def getMomentumReference(self): return self.momentumReference
    def setAlphaGravity(self, a): self.alphaGravity = float(a)
    def getAlphaGravity(self): return self.alphaGravity


# =================================================================================================
#  Costs (MPC/src/variableSamplingMPC/costsVSMPC.cpp); each owns a dense nVar x nVar Hessian
#  (UT/src/IQPCost.cpp:16-20)
# =================================================================================================
class _Cost:
    def __init__(self, nVar):
        self.nVar = nVar
        self.firstUpdate = True

    def configureSizeHessianAndGradient(self):
        self.hessian = np.zeros((self.nVar, self.nVar))
        self.gradient = np.zeros(self.nVar)


def _W_of_rpy(rpy):
    """costsVSMPC.cpp:276-282 / systemDynamicsVSMPC.cpp:133-139."""
    W = np.zeros((3, 3))
    W[0, 0] = 1.0
    W[1, 1] = math.cos(rpy[0])
    W[2, 1] = -math.sin(rpy[0])
    W[0, 2] = -math.sin(rpy[1])
    W[1, 2] = math.cos(rpy[1]) * math.sin(rpy[0])
    W[2, 2] = math.cos(rpy[0]) * math.cos(rpy[1])
    return W


def _locked_inertia(robot: RobotData):
    """(X^T M_b X).block(3,3,3,3), X = Ad(G_H_B) (systemDynamicsVSMPC.cpp:110-130, costsVSMPC.cpp:268-285)."""
    r = robot.p_com - robot.base_pos
    X = adjoint_transform(robot.wRb, r)
    return (X.T @ robot.mass_matrix_base @ X)[3:6, 3:6]


class ReferenceTrackingCost(_Cost):
    """costsVSMPC.cpp:5-307."""

    def __init__(self, nVar, nStates, nIter):
        super().__init__(nVar)
        self.nStates, self.nIter = nStates, nIter

    def readConfigParameters(self, p, qpInput, trajectories):
        self.wCoMPos = np.array(p["weightCoMPos"], float)
        self.wCoMPosError = np.array(p["weightCoMPosError"], float)
        self.wLinMom = np.array(p["weightLinMom"], float)
        self.wRPY = np.array(p["weightRPY"], float)
        self.wRPYError = np.array(p["weightRPYError"], float)
        self.wAngMom = np.array(p["weightAngMom"], float)
        self.nIterSmall = p["nIterSmall"]
        t = trajectories["POSITION_TRAJECTORY"]
        self.trajManager = TrajectoryManager(t["arrays"], t["fps"], 1 / p["periodMPCLargeSteps"])  # :68
        self.ratio = int(round(p["periodMPCLargeSteps"] / p["periodMPCSmallSteps"]))  # :69

    def configureDynVectorsSize(self, qpInput):  # :74-119
        self.robot = qpInput.getRobot()
        nS = self.nStates
        Q = np.zeros((nS, nS))
        Q[0:3, 0:3] = np.diag(self.wCoMPos)
        Q[3:6, 3:6] = np.diag(self.wLinMom)
        Q[6:9, 6:9] = np.diag(self.wRPY)
        Q[9:12, 9:12] = np.diag(self.wAngMom)
        Q[20:23, 20:23] = np.diag(self.wCoMPosError)
        Q[23:26, 23:26] = np.diag(self.wRPYError)
        self.Q = Q
        self.stateReference = np.zeros((nS, self.nIter))
        nc = self.nIter - self.nIterSmall + 1
        self.initialCoMPos = self.robot.p_com.copy()
        self.initialRPY = self.robot.rpy.copy()
        self.updateInertiaMatrix()
        self.posRef = np.zeros((3, nc))
        self.linMomRef = np.zeros((3, nc))
        self.rpyRef = np.zeros((3, nc))
        self.angMomRef = np.zeros((3, nc))
        for i in range(nc):
            self.posRef[:, i] = self.initialCoMPos + self.trajManager.getCurrentValue("positionCoM")
            self.linMomRef[:, i] = self.robot.wRb.T @ (self.robot.getTotalMass()
                                                       * self.trajManager.getCurrentValue("velocityCoM"))
            self.rpyRef[:, i] = self.initialRPY + self.trajManager.getCurrentValue("RPY")
            self.angMomRef[:, i] = self.inertia @ self.W @ self.trajManager.getCurrentValue("RPYDot")
        self._push_refs()
        self.counter = self.ratio - 1

    def _push_refs(self):  # set*Reference, :183-264
        for rows, ref in ((CoMPosIdx, self.posRef), (linMomIdx, self.linMomRef),
                          (rpyIdx, self.rpyRef), (angMomIdx, self.angMomRef)):
            for i in range(self.nIter):
                col = 0 if i < self.nIterSmall else i - self.nIterSmall
                self.stateReference[rows[0]:rows[0] + 3, i] = ref[:, col]

    def updateInertiaMatrix(self):  # :266-286
        self.W = _W_of_rpy(self.robot.rpy)
        self.inertia = _locked_inertia(self.robot)

    def computeHessianAndGradient(self, qpInput):  # :121-181
        self.gradient[:] = 0.0
        if self.counter == self.ratio - 1:
            self.trajManager.advanceTrajectory()
            tm = self.trajManager
            self.posRef = np.hstack([self.posRef[:, 1:],
                                     (self.initialCoMPos + tm.getCurrentValue("positionCoM"))[:, None]])
            self.linMomRef = np.hstack([self.linMomRef[:, 1:],
                                        (self.robot.wRb.T @ (self.robot.getTotalMass()
                                                             * tm.getCurrentValue("velocityCoM")))[:, None]])
            self.rpyRef = np.hstack([self.rpyRef[:, 1:],
                                     (self.initialRPY + tm.getCurrentValue("RPY"))[:, None]])
            self.updateInertiaMatrix()
            self.angMomRef = np.hstack([self.angMomRef[:, 1:],
                                        (self.inertia @ self.W @ tm.getCurrentValue("RPYDot"))[:, None]])
            self._push_refs()
            qpInput.setPosCoMReference(self.posRef[:, 0])
            qpInput.setRPYReference(self.rpyRef[:, 0])
            qpInput.setMomentumReference(np.concatenate([self.linMomRef[:, 0], self.angMomRef[:, 0]]))
            self.counter = 0
        else:
            self.counter += 1
        nS = self.nStates
        if self.firstUpdate:
            self.hessian[:] = 0.0
            for i in range(1, self.nIter + 1):
                self.hessian[i * nS:(i + 1) * nS, i * nS:(i + 1) * nS] = self.Q
            self.firstUpdate = False
        for i in range(1, self.nIter + 1):
            self.gradient[i * nS:(i + 1) * nS] = -self.Q @ self.stateReference[:, i - 1]
        return True


class RegualarizationCost(_Cost):
    """costsVSMPC.cpp:309-425 (class name spelled as in the reference)."""

    def __init__(self, nVar, nStates, nJoints, nThrottle):
        super().__init__(nVar)
        self.nStates, self.nCtrlJoints, self.nJets = nStates, nJoints, nThrottle

    def readConfigParameters(self, p, qpInput, trajectories):
        self.nIter = p["nIter"]
        self.nSmallSteps = p["nIterSmall"]
        self.ctrlHorizon = p["controlHorizon"]
        self.weightDeltaJoint = np.diag(np.array(p["weightDeltaJoint"], float))
        self.weightThrottleMatrix = p["weightThrottle"] * np.eye(self.nJets)

    def configureDynVectorsSize(self, qpInput):
        pass

    def computeHessianAndGradient(self, qpInput):  # :369-413
        if self.firstUpdate:
            self.hessian[:] = 0.0
            self.gradient[:] = 0.0
            nS, nJ, nT = self.nStates, self.nCtrlJoints, self.nJets
            base = nS * (self.nIter + 1)
            for i in range(self.ctrlHorizon):
                o = base + i * nJ
                self.hessian[o:o + nJ, o:o + nJ] = self.weightDeltaJoint
            tb = base + self.ctrlHorizon * nJ
            for i in range(self.ctrlHorizon - self.nSmallSteps):
                a, b = tb + i * nT, tb + (i + 1) * nT
                self.hessian[a:a + nT, a:a + nT] += self.weightThrottleMatrix
                self.hessian[b:b + nT, a:a + nT] -= self.weightThrottleMatrix
                self.hessian[a:a + nT, b:b + nT] -= self.weightThrottleMatrix
                self.hessian[b:b + nT, b:b + nT] += self.weightThrottleMatrix
            self.firstUpdate = False
        return True


class ThrottleInitialValueCost(_Cost):
    """costsVSMPC.cpp:427-499."""

    def __init__(self, nVar, nStates, nJoints, nThrottle):
        super().__init__(nVar)
        self.nStates, self.nCtrlJoints, self.nJets = nStates, nJoints, nThrottle

    def readConfigParameters(self, p, qpInput, trajectories):
        self.nIter = p["nIter"]
        self.weightThrottle = p["weightInitialThrottle"]
        self.ctrlHorizon = p["controlHorizon"]

    def configureDynVectorsSize(self, qpInput):
        self.jetModel = qpInput.getJetModel()

    def computeHessianAndGradient(self, qpInput):  # :468-487
        o = self.nStates * (self.nIter + 1) + self.nCtrlJoints * self.ctrlHorizon
        if self.firstUpdate:
            self.hessian[o:o + self.nJets, o:o + self.nJets] = self.weightThrottle * np.eye(self.nJets)
            self.firstUpdate = False
        jm = self.jetModel
        for i in range(self.nJets):
            self.gradient[o + i] = -self.weightThrottle * jm.compute_v(
                jm.standardizeThrottle_u2T(qpInput.getThrottleMPC()[i]))
        return True


class JointPositionRegularizationCost(_Cost):
    """costsVSMPC.cpp:501-603."""

    def __init__(self, nVar, nStates, nJoints):
        super().__init__(nVar)
        self.nStates, self.nCtrlJoints = nStates, nJoints

    def readConfigParameters(self, p, qpInput, trajectories):  # :511-552
        self.nIter = p["nIter"]
        self.weightJointPos = p["weightRegularizationJointPos"]
        self.ctrlHorizon = p["controlHorizon"]
        self.ctrlJointsNames = list(p["controlledJoints"])
        self.robot = qpInput.getRobot()
        self.jointPosReference = np.zeros(self.nCtrlJoints)
        for i in range(self.nCtrlJoints):
            for j in range(self.robot.getNJoints()):
                if self.ctrlJointsNames[i] == self.robot.joint_names[j]:
                    self.jointPosReference[i] = self.robot.joint_pos[j]
                    break

    def configureDynVectorsSize(self, qpInput):
        pass

    def computeHessianAndGradient(self, qpInput):  # :558-592
        nS, nJ = self.nStates, self.nCtrlJoints
        base = nS * (self.nIter + 1)
        if self.firstUpdate:
            self.hessian[:] = 0.0
            self.gradient[:] = 0.0
            for i in range(self.ctrlHorizon):
                o = base + i * nJ
                self.hessian[o:o + nJ, o:o + nJ] = self.weightJointPos * np.eye(nJ)
            self.firstUpdate = False
        jointPos = np.zeros(nJ)
        for i in range(nJ):
            for j in range(self.robot.getNJoints()):
                if self.ctrlJointsNames[i] == self.robot.joint_names[j]:
                    jointPos[i] = qpInput.getOutputQPJointsPosition()[j]
                    break
        for i in range(self.ctrlHorizon):
            o = base + i * nJ
            self.gradient[o:o + nJ] = self.weightJointPos * (jointPos - self.jointPosReference)
        return True


# =================================================================================================
#  Dynamics (MPC/src/variableSamplingMPC/systemDynamicsVSMPC.cpp)
# =================================================================================================
class _Dyn:
    def __init__(self, nStates, nJoints, nThrottle):
        self.A = np.zeros((nStates, nStates))
        self.BJ = np.zeros((nStates, nJoints))
        self.BT = np.zeros((nStates, nThrottle))
        self.c = np.zeros(nStates)

    def _zero(self):
        self.A[:] = 0
        self.BJ[:] = 0
        self.BT[:] = 0
        self.c[:] = 0


class AngularMomentumDynamicVS(_Dyn):
    """:6-226 (jointsLambdaOption 'unfiltered' and 'constant')."""

    def configure(self, p, qpInput, trajectories):
        self.option = p["jointsLambdaOption"]
        assert self.option in ("unfiltered", "constant")
        self.robot = qpInput.getRobot()
        self.robotReference = qpInput.getRobotReference()
        r = self.robot
        self.relJacobianInit = r.J_rel_body.copy()
        self.matJetAxisInit = r.jet_axes.copy()
        self.matJetArmsInit = r.jet_arms.copy()
        self.sel = [j for name in p["controlledJoints"]
                    for j in range(r.getNJoints()) if name == r.joint_names[j]]
        self.rpyInit = self.robotReference.rpy.copy()  # :67

    def updateInitialStates(self, qpInput):
        self.updateRPY()
        self.computeLambdaAng(qpInput)
        return self.computeAngularMomentumMatrices()

    def updateRPY(self):  # :105-157 (the unused M_bs/J_s/G_T_B/M_G block :117-126 is dead code)
        rr = self.robotReference
        self.wRb = rr.wRb
        self.B_omega_B = self.wRb.T @ self.robot.omega_world
        self.inertia = _locked_inertia(rr)
        rpy = rr.rpy
        Wi = np.zeros((3, 3))
        Wi[0, 0] = 1.0
        Wi[0, 1] = math.sin(rpy[0]) * math.tan(rpy[1])
        Wi[1, 1] = math.cos(rpy[0])
        Wi[2, 1] = math.sin(rpy[0]) / math.cos(rpy[1])
        Wi[0, 2] = math.cos(rpy[0]) * math.tan(rpy[1])
        Wi[1, 2] = -math.sin(rpy[0])
        Wi[2, 2] = math.cos(rpy[0]) / math.cos(rpy[1])
        self.WInverse = Wi

    def getRelativeJacobianCoM(self, i):  # :208-226
        rr = self.robotReference
        jac = rr.J_jet_lin[i] - rr.J_com
        return rr.wRb.T @ jac

    def computeLambdaAng(self, qpInput):  # :159-206
        rr = self.robotReference
        nJ = rr.getNJoints()
        lam = np.zeros((3, nJ))
        Rt = self.wRb.T
        if self.option == "unfiltered":
            for i in range(rr.getNJets()):
                T = rr.jet_thrusts[i]
                Sa = from_vec_to_skew(Rt @ rr.jet_axes[i])
                lam -= T * Sa @ self.getRelativeJacobianCoM(i)
                lam -= T * from_vec_to_skew(Rt @ rr.jet_arms[i]) @ Sa @ rr.J_rel_body[i][3:6, :]
        else:
            for i in range(rr.getNJets()):
                Si = np.zeros((3, 6))
                Sa = from_vec_to_skew(Rt @ self.matJetAxisInit[i])
                Si[:, 0:3] = Sa
                Si[:, 3:6] = from_vec_to_skew(Rt @ self.matJetArmsInit[i]) @ Sa
                Si *= self.robot.jet_thrusts[i]
                lam -= Si @ self.relJacobianInit[i]
        self.lambdaAngB = lam[:, self.sel]

    def computeAngularMomentumMatrices(self):  # :79-103
        self._zero()
        self.A[6:9, 9:12] = self.WInverse @ np.linalg.inv(self.inertia)
        self.A[9:12, 9:12] -= from_vec_to_skew(self.B_omega_B)
        self.A[9:12, 12:16] = self.robotReference.A_mom_body[3:6, :]
        self.BJ[9:12, 0:N_JOINTS] = self.lambdaAngB
        self.A[23:26, 6:9] = np.eye(3)
        self.c[23:26] = -self.rpyInit
        return True


class LinearMomentumDynamicVS(_Dyn):
    """:228-350."""

    def configure(self, p, qpInput, trajectories):
        self.option = p["jointsLambdaOption"]
        t = trajectories["TRAJECTORY_MANAGER"]
        self.trajectoryManager = TrajectoryManager(t["arrays"], t["fps"], 1 / p["periodMPC"])  # :272
        self.robot = qpInput.getRobot()
        self.robotReference = qpInput.getRobotReference()
        self.relJacobianInit = self.robot.J_rel_body.copy()
        self.matJetAxisInit = self.robot.jet_axes.copy()

    def updateInitialStates(self, qpInput):
        self.computeLambdaLin(qpInput)
        return self.computeLinearMomentumMatrices(qpInput)

    def computeLambdaLin(self, qpInput):  # :321-350
        rr = self.robotReference
        lam = np.zeros((3, rr.getNJoints()))
        self.wRb = rr.wRb
        self.B_omega_B = self.wRb.T @ self.robot.omega_world
        for i in range(rr.getNJets()):
            if self.option == "constant":
                lam -= rr.jet_thrusts[i] * from_vec_to_skew(self.wRb.T @ self.matJetAxisInit[i]) \
                    @ self.relJacobianInit[i][3:6, :]
            else:
                lam -= rr.jet_thrusts[i] * from_vec_to_skew(self.wRb.T @ rr.jet_axes[i]) \
                    @ rr.J_rel_body[i][3:6, :]
        self.lambdaB = lam[:, 3:3 + N_JOINTS]  # middleCols(3, 8), hard-coded (:348)
        return True

    def computeLinearMomentumMatrices(self, qpInput):  # :288-319
        self._zero()
        rr = self.robotReference
        self.A[0:3, 3:6] = 1 / rr.getTotalMass() * self.wRb
        self.A[3:6, 3:6] -= from_vec_to_skew(self.B_omega_B)
        self.A[3:6, 12:16] = rr.A_mom_body[0:3, :]
        self.BJ[3:6, 0:N_JOINTS] = self.lambdaB
        alpha = self.trajectoryManager.getCurrentValue("alphaGravity")[0]
        self.c[3:6] = alpha * rr.getTotalMass() * (self.wRb.T @ rr.gravity)
        qpInput.setAlphaGravity(alpha)
        self.trajectoryManager.advanceTrajectory()
        self.A[20:23, 0:3] = np.eye(3)
        self.c[20:23] = -qpInput.getPosCoMReference()
        return True


class JetDynamicVS(_Dyn):
    """:352-461."""

    def configure(self, p, qpInput, trajectories):
        self.useJetDynamic = p["useJetDynamic"]
        self.useEstimatedThrust = p["useEstimatedThrust"]
        self.robot = qpInput.getRobot()
        self.jetModel = qpInput.getJetModel()

    def computeF(self, T, Td):
        jm = self.jetModel
        return jm.compute_f(jm.standardizeThrust_u2T(T), jm.standardizeThrustDot_u2T(Td)) \
            * jm.getThrustStandardDeviation_u2T()

    def computeG(self, T, Td):
        jm = self.jetModel
        return jm.compute_g(jm.standardizeThrust_u2T(T), jm.standardizeThrustDot_u2T(Td)) \
            * jm.getThrustStandardDeviation_u2T()

    def compute_dh_dT(self, T, Td, throttle):
        jm = self.jetModel
        T, Td = jm.standardizeThrust_u2T(T), jm.standardizeThrustDot_u2T(Td)
        u = jm.standardizeThrottle_u2T(throttle)
        return jm.compute_df_dT(T, Td) + jm.compute_dg_dT(T, Td) * jm.compute_v(u)

    def compute_dh_dTDot(self, T, Td, throttle):
        jm = self.jetModel
        T, Td = jm.standardizeThrust_u2T(T), jm.standardizeThrustDot_u2T(Td)
        u = jm.standardizeThrottle_u2T(throttle)
        return jm.compute_df_dTdot(T, Td) + jm.compute_dg_dTdot(T, Td) * jm.compute_v(u)

    def updateInitialStates(self, qpInput):  # :384-429
        self._zero()
        if self.useJetDynamic:
            self.A[12:16, 16:20] = np.eye(4)
            for i in range(self.robot.getNJets()):
                if self.useEstimatedThrust:
                    T = self.robot.jet_thrusts[i]
                    Td = qpInput.getEstimatedThrustDot()[i]
                else:
                    T = qpInput.getThrustDesMPC()[i]
                    Td = qpInput.getThrustDotDesMPC()[i]
                u = qpInput.getThrottleMPC()[i]
                self.A[16 + i, 12 + i] = self.compute_dh_dT(T, Td, u)
                self.A[16 + i, 16 + i] += self.compute_dh_dTDot(T, Td, u)
                self.BT[16 + i, i] = self.computeG(qpInput.getThrustDesMPC()[i],
                                                   qpInput.getThrustDotDesMPC()[i])
                self.c[16 + i] = self.computeF(T, Td) - self.compute_dh_dT(T, Td, u) * T \
                    - self.compute_dh_dTDot(T, Td, u) * Td
        else:
            self.BT[12:16, 0:4] = np.eye(4)
        return True


class SystemDynamicVS:
    """:463-585.  Sub-models are created Angular, Linear, Jet in that order (:478-482)."""

    def __init__(self, nStates, nJoints, nThrottle):
        self.dims = (nStates, nJoints, nThrottle)
        self.vectorDynamic = []

    def configure(self, p, qpInput, trajectories):
        self.vectorDynamic = [AngularMomentumDynamicVS(*self.dims), LinearMomentumDynamicVS(*self.dims),
                              JetDynamicVS(*self.dims)]
        for d in self.vectorDynamic:
            d.configure(p, qpInput, trajectories)
        return True

    def updateDynamicMatrices(self, qpInput):
        for d in self.vectorDynamic:
            d.updateInitialStates(qpInput)
        return True

    def getAMatrix(self): return sum(d.A for d in self.vectorDynamic)
    def getBJointsMatrix(self): return sum(d.BJ for d in self.vectorDynamic)
    def getBThrottleMatrix(self): return sum(d.BT for d in self.vectorDynamic)
    def getCVector(self): return sum(d.c for d in self.vectorDynamic)


# =================================================================================================
#  Constraints (constraintsVSMPC.cpp); each owns a dense nCon x nVar block (UT/src/IQPConstraint.cpp:60-65)
# =================================================================================================
class _Constraint:
    def __init__(self, nVar, nCon):
        self.nVar, self.nConstraints = nVar, nCon

    def configureSizeConstraintMatrixAndBounds(self):
        self.linearMatrix = np.zeros((self.nConstraints, self.nVar))
        self.lowerBound = np.zeros(self.nConstraints)
        self.upperBound = np.zeros(self.nConstraints)

    def getNConstraints(self):
        return self.nConstraints


def time_grid(p) -> np.ndarray:
    """dt_k of constraintsVSMPC.cpp:45-51,76-84,156-159."""
    nS, dS, dL = p["nIterSmall"], p["periodMPCSmallSteps"], p["periodMPCLargeSteps"]
    beta2 = (dL - nS * dS) / (nS * (nS - 1))
    beta1 = dS - beta2
    warp = lambda t: beta1 * t + beta2 * t * t
    return np.array([warp(i + 1) - warp(i) if i < nS else dL for i in range(p["nIter"])])


class ConstraintSystemDynamicVS(_Constraint):
    """:5-159."""

    def __init__(self, nVar, nStates, nJoints, nThrottle, nIter):
        super().__init__(nVar, nStates * nIter)
        self.systemDynamicVS = SystemDynamicVS(nStates, nJoints, nThrottle)
        self.nStates, self.nJoints, self.nThrottle, self.nIter = nStates, nJoints, nThrottle, nIter

    def readConfigParameters(self, p, qpInput, trajectories):
        self.nSmallSteps = p["nIterSmall"]
        self.ctrlHorizon = p["controlHorizon"]
        self.deltaTSmallSteps = p["periodMPCSmallSteps"]
        self.deltaTLargeSteps = p["periodMPCLargeSteps"]
        self.beta2 = (self.deltaTLargeSteps - self.nSmallSteps * self.deltaTSmallSteps) \
            / (self.nSmallSteps * (self.nSmallSteps - 1))
        self.beta1 = self.deltaTSmallSteps - self.beta2
        return self.systemDynamicVS.configure(p, qpInput, trajectories)

    def configureDynVectorsSize(self, qpInput):
        pass

    def warp_function(self, t):
        return self.beta1 * t + self.beta2 * t * t

    def computeConstraintsMatrixAndBounds(self, qpInput):  # :61-142
        sd = self.systemDynamicVS
        sd.updateDynamicMatrices(qpInput)
        A, BJ, BT, c = sd.getAMatrix(), sd.getBJointsMatrix(), sd.getBThrottleMatrix(), sd.getCVector()
        self.A, self.BJ, self.BT, self.c = A, BJ, BT, c
        nS, nJ, nT, N = self.nStates, self.nJoints, self.nThrottle, self.nIter
        M = self.linearMatrix
        M[:] = 0
        self.lowerBound[:] = 0
        self.upperBound[:] = 0
        base = nS * (N + 1)
        self.dt = np.zeros(N)
        for i in range(N):
            if i < self.nSmallSteps:
                dT = self.warp_function(i + 1) - self.warp_function(i)
            else:
                dT = self.deltaTLargeSteps
            self.dt[i] = dT
            r = slice(i * nS, (i + 1) * nS)
            M[r, i * nS:(i + 1) * nS] = np.eye(nS) + dT * A
            M[r, (i + 1) * nS:(i + 2) * nS] = -np.eye(nS)
            jb = i if i < self.ctrlHorizon else self.ctrlHorizon - 1
            M[r, base + jb * nJ: base + (jb + 1) * nJ] = dT * BJ
            if i < self.nSmallSteps:
                tb = 0
            elif i < self.ctrlHorizon:
                tb = i - (self.nSmallSteps - 1)
            else:
                tb = self.ctrlHorizon - self.nSmallSteps
            o = base + self.ctrlHorizon * nJ + tb * nT
            M[r, o:o + nT] = dT * BT
            self.lowerBound[r] = -dT * c
            self.upperBound[r] = -dT * c
        return True


class ConstraintInitialState(_Constraint):
    """constraintsVSMPC.cpp:161-277 + IQPUtilsMPC.cpp:57-92."""

    def __init__(self, nStates, nVar):
        super().__init__(nVar, nStates)
        self.nStates = nStates

    def readConfigParameters(self, p, qpInput, trajectories):
        self.useEstimatedThrust = p["useEstimatedThrust"]

    def configureDynVectorsSize(self, qpInput):  # :184-204
        self.robot = qpInput.getRobot()
        self.rpyOld = self.robot.rpy.copy()
        self.nTurns = np.zeros(3)

    def unwrapRPY(self):  # :232-247
        rpy = self.robot.rpy
        for i in range(3):
            if rpy[i] - self.rpyOld[i] > math.pi:
                self.nTurns[i] -= 1
            elif rpy[i] - self.rpyOld[i] < -math.pi:
                self.nTurns[i] += 1
        self.rpyUnwrapped = rpy + 2 * math.pi * self.nTurns
        self.rpyOld = rpy.copy()

    def updateInitialState(self, qpInput):  # :206-230
        self.unwrapRPY()
        x = np.zeros(self.nStates)
        r = self.robot
        x[0:3] = r.p_com
        x[3:6] = r.momentum_body[0:3]
        x[6:9] = self.rpyUnwrapped
        x[9:12] = r.momentum_body[3:6]
        if self.useEstimatedThrust:
            x[12:16] = r.jet_thrusts
            x[16:20] = qpInput.getEstimatedThrustDot()
        else:
            x[12:16] = qpInput.getThrustDesMPC()
            x[16:20] = qpInput.getThrustDotDesMPC()
        x[20:23] = r.p_com - qpInput.getPosCoMReference()
        x[23:26] = self.rpyUnwrapped - qpInput.getRPYReference()
        self.initialState = x
        return True

    def computeConstraintsMatrixAndBounds(self, qpInput):  # IQPUtilsMPC.cpp:71-92
        self.linearMatrix[:] = 0
        self.updateInitialState(qpInput)
        self.linearMatrix[0:self.nStates, 0:self.nStates] = np.eye(self.nStates)
        self.lowerBound[:] = self.initialState
        self.upperBound[:] = self.initialState
        return True


class ThrottleConstraint(_Constraint):
    """constraintsVSMPC.cpp:279-386 (sized with nIter, filled with controlHorizon: :283 vs :343)."""

    def __init__(self, nVar, nStates, nIter, nSmallSteps):
        super().__init__(nVar, N_THRUSTS * (nIter - nSmallSteps + 1))
        self.nIter, self.nSmallSteps, self.nStates = nIter, nSmallSteps, nStates

    def readConfigParameters(self, p, qpInput, trajectories):
        self.throttleMaxValue = p["throttleMax"]
        self.throttleMinValue = p["throttleMin"]
        self.ctrlHorizon = p["controlHorizon"]
        self.ratio = int(round(p["periodMPCLargeSteps"] / p["periodMPCSmallSteps"]))

    def configureDynVectorsSize(self, qpInput):  # :326-336
        jm = self.jetModel = qpInput.getJetModel()
        self.vMax = jm.compute_v(jm.standardizeThrottle_u2T(self.throttleMaxValue))
        self.vMin = jm.compute_v(jm.standardizeThrottle_u2T(self.throttleMinValue))
        self.nJoints, self.nThrottle = N_JOINTS, N_THRUSTS
        self.counter = self.ratio - 1

    def computeConstraintsMatrixAndBounds(self, qpInput):  # :338-374
        self.linearMatrix[:] = 0
        self.lowerBound[:] = 0
        self.upperBound[:] = 0
        nT = self.nThrottle
        base = self.nStates * (self.nIter + 1) + self.nJoints * self.ctrlHorizon
        jm = self.jetModel
        for i in range(self.ctrlHorizon - self.nSmallSteps + 1):
            self.linearMatrix[i * nT:(i + 1) * nT, base + i * nT: base + (i + 1) * nT] = np.eye(nT)
            if self.counter != self.ratio - 1 and i == 0:
                for j in range(N_THRUSTS):
                    v = jm.compute_v(jm.standardizeThrottle_u2T(qpInput.getThrottleMPC()[j]))
                    self.lowerBound[j] = v
                    self.upperBound[j] = v
            else:
                self.lowerBound[i * nT:(i + 1) * nT] = self.vMin
                self.upperBound[i * nT:(i + 1) * nT] = self.vMax
        if self.counter == self.ratio - 1:
            self.counter = 0
        else:
            self.counter += 1
        return True


class JointPositionConstraint(_Constraint):
    """constraintsVSMPC.cpp:388-468 — present in the reference but NOT registered in the shipped problem
    (variableSamplingMPC.cpp:77-84 lists three constraints) and its XML has no jointPos_max / jointPos_min; offered here as the
    optional per-instance joint-limit rows of BASELINE configs[4] (SURVEY §8d Config 5), registered after the throttle rows
    when the parameters carry `jointPos_max` / `jointPos_min` [degrees, :421-423].

    Semantics kept: the class is sized nJoints * nIter rows (:391), block i < controlHorizon bounds dq_i — the displacement of
    the controlled joints from the commanded posture that acts on knot i — by
        jointPos_min - q_cmd[controlled] <= dq_i <= jointPos_max - q_cmd[controlled]              (:450-453)
    for EVERY block (per block, not cumulative: dq_i is a displacement, the accumulation over ticks happens in q_cmd).
    Deliberately NOT ported: the `m_firstIteriation = false` inside the loop (:440-449), which leaves the identity rows of
    blocks 1.. unset and turns their bounds into constraints on all-zero rows (infeasible whenever q_cmd leaves the limits)."""

    def __init__(self, nVar, nStates, nJoints, nIter):
        super().__init__(nVar, nJoints * nIter)
        self.nIter, self.nStates, self.nJoints = nIter, nStates, nJoints

    def readConfigParameters(self, p, qpInput, trajectories):
        self.jointPositionMax = np.asarray(p["jointPos_max"], float) * math.pi / 180.0
        self.jointPositionMin = np.asarray(p["jointPos_min"], float) * math.pi / 180.0
        assert self.jointPositionMax.shape == (self.nJoints,) and self.jointPositionMin.shape == (self.nJoints,)
        self.ctrlHorizon = p["controlHorizon"]
        self.controlledJoints = list(p["controlledJoints"])

    def configureDynVectorsSize(self, qpInput):  # :428-433
        self.robot = qpInput.getRobot()
        # the reference indexes getOutputQPJointsPosition().segment(3, nJoints) (:451): the controlled joints of the shipped
        # robot; here by name, like every other class of the problem
        self.sel = [i for name in self.controlledJoints for i in range(self.robot.getNJoints())
                    if name == self.robot.joint_names[i]]

    def computeConstraintsMatrixAndBounds(self, qpInput):  # :434-456
        self.linearMatrix[:] = 0
        self.lowerBound[:] = 0
        self.upperBound[:] = 0
        nJ = self.nJoints
        base = self.nStates * (self.nIter + 1)
        qcmd = np.asarray(qpInput.getOutputQPJointsPosition(), float)[self.sel]
        for i in range(self.ctrlHorizon):
            self.linearMatrix[i * nJ:(i + 1) * nJ, base + i * nJ: base + (i + 1) * nJ] = np.eye(nJ)
            self.lowerBound[i * nJ:(i + 1) * nJ] = self.jointPositionMin - qcmd
            self.upperBound[i * nJ:(i + 1) * nJ] = self.jointPositionMax - qcmd
        return True


# =================================================================================================
#  Exact QP solve: the arbiter at "matched KKT tolerance" (SURVEY §8c).
# =================================================================================================
def solve_qp_exact(P, q, A, l, u, tol=1e-9, max_iter=100):
    """Exact minimiser of 1/2 z'Pz + q'z s.t. l <= Az <= u by a primal-dual active-set iteration on
    the dense KKT system; returns (z, y, info).  The problem class here is strictly
    convex on the equality-feasible subspace, so the minimiser is unique; the returned ``info``
    holds the KKT residuals certifying it, independent of how the active set was found.
    """
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = P.shape[0]
    Psp = sp.csc_matrix(P)
    Asp = sp.csr_matrix(A)
    nz = np.abs(A).sum(axis=1) > 0
    if np.any(~nz & ((l > 0) | (u < 0))):
        raise ValueError("infeasible all-zero row")
    eq = nz & (l == u)
    ineq = np.where(nz & ~eq)[0]
    eq_idx = np.where(eq)[0]
    act_lo = np.zeros(len(ineq), bool)
    act_up = np.zeros(len(ineq), bool)
    seen = set()
    z = None
    for it in range(max_iter):
        rows = np.concatenate([eq_idx, ineq[act_lo], ineq[act_up]]).astype(int)
        rhs_c = np.concatenate([l[eq_idx], l[ineq[act_lo]], u[ineq[act_up]]])
        Aw = Asp[rows]
        m = len(rows)
        K = sp.bmat([[Psp, Aw.T], [Aw, None]], format="csc") if m else Psp.tocsc()
        lu = spla.splu(K)
        rhs = np.concatenate([-q, rhs_c])
        sol = lu.solve(rhs)
        for _ in range(2):  # iterative refinement (KKT is indefinite; SuperLU pivots)
            sol = sol + lu.solve(rhs - K @ sol)
        z, lam = sol[:n], sol[n:]
        ne = len(eq_idx)
        nlo = int(act_lo.sum())
        lam_lo = lam[ne:ne + nlo]          # multiplier of an active lower bound must be <= 0
        lam_up = lam[ne + nlo:]            # multiplier of an active upper bound must be >= 0
        Az = A[ineq] @ z
        new_lo = act_lo.copy()
        new_up = act_up.copy()
        new_lo[np.where(act_lo)[0][lam_lo > tol]] = False
        new_up[np.where(act_up)[0][lam_up < -tol]] = False
        viol_lo = (~act_lo & ~act_up) & (Az < l[ineq] - tol)
        viol_up = (~act_lo & ~act_up) & (Az > u[ineq] + tol)
        new_lo |= viol_lo
        new_up |= viol_up
        if np.array_equal(new_lo, act_lo) and np.array_equal(new_up, act_up):
            break
        key = (new_lo.tobytes(), new_up.tobytes())
        if key in seen:  # cycling guard: add only the single most violated / drop the worst sign
            new_lo, new_up = act_lo.copy(), act_up.copy()
            cand = []
            if lam_lo.size and lam_lo.max() > tol:
                cand.append((lam_lo.max(), "dl", np.where(act_lo)[0][lam_lo.argmax()]))
            if lam_up.size and (-lam_up).max() > tol:
                cand.append(((-lam_up).max(), "du", np.where(act_up)[0][(-lam_up).argmax()]))
            if cand:
                _, kind, j = max(cand)
                (new_lo if kind == "dl" else new_up)[j] = False
            else:
                v = np.where(viol_lo, l[ineq] - Az, 0) + np.where(viol_up, Az - u[ineq], 0)
                j = int(v.argmax())
                (new_lo if viol_lo[j] else new_up)[j] = True
        seen.add(key)
        act_lo, act_up = new_lo, new_up
    else:
        raise RuntimeError("active-set iteration did not converge")
    y = np.zeros(A.shape[0])
    y[rows] = lam
    Az_full = A @ z
    info = dict(
        iters=it + 1,
        n_active=int(act_lo.sum() + act_up.sum()),
        stationarity=float(np.abs(P @ z + q + A.T @ y).max()),
        primal=float(max(np.maximum(l - Az_full, 0).max(), np.maximum(Az_full - u, 0).max())),
        dual_sign=float(max(np.maximum(y[ineq[act_lo]], 0).max(initial=0.0),
                            np.maximum(-y[ineq[act_up]], 0).max(initial=0.0))),
    )
    return z, y, info


# =================================================================================================
class IMPCProblem:
    """MPC/src/IMPCProblem/IMPCProblem.cpp:3-298 with the OsqpEigen call replaced by a pluggable
    ``qp_solver(P,q,A,l,u, warm) -> (z, status)`` (default: the exact solver above)."""

    SOLVED = 1  # OsqpEigen::Status::Solved

    def __init__(self, qp_solver=None):
        self.vectorCosts = []
        self.vectorConstraints = []
        self.firstUpdate = True
        self.qp_solver = qp_solver
        self.statusQPProblem = None

    def configure(self, params, qpInput, trajectories, phase0=0):  # :3-148
        """phase0 (NOT in the reference; mirrors the phase0 argument of vsmpc_configure): number of ticks by which the two
        20-tick counters (ReferenceTrackingCost::m_counter, ThrottleConstraint::m_counter) start ahead, so that a batch can
        be checked with staggered throttle-release phases.  0 = reference behaviour."""
        self.setCostAndConstraints(params, qpInput)
        for c in self.vectorCosts:
            c.readConfigParameters(params, qpInput, trajectories)
        for c in self.vectorConstraints:
            c.readConfigParameters(params, qpInput, trajectories)
        self.hessian = np.zeros((self.nVar, self.nVar))
        self.gradient = np.zeros(self.nVar)
        for c in self.vectorCosts:
            c.configureDynVectorsSize(qpInput)
            c.configureSizeHessianAndGradient()
            if phase0 and hasattr(c, "counter") and hasattr(c, "ratio"):
                c.counter = (c.ratio - 1 + int(phase0)) % c.ratio
        for c in self.vectorCosts:
            c.computeHessianAndGradient(qpInput)
            self.hessian += c.hessian
            self.gradient += c.gradient
        n = 0
        for c in self.vectorConstraints:
            c.configureDynVectorsSize(qpInput)
            c.configureSizeConstraintMatrixAndBounds()
            if phase0 and hasattr(c, "counter") and hasattr(c, "ratio"):
                c.counter = (c.ratio - 1 + int(phase0)) % c.ratio
            n += c.getNConstraints()
        self.linearMatrix = np.zeros((n, self.nVar))
        self.lowerBound = np.zeros(n)
        self.upperBound = np.zeros(n)
        self._gather_constraints(qpInput)
        self.nConstraints = n
        self.outputQP = np.zeros(self.nVar)
        return True

    def _gather_constraints(self, qpInput):
        n = 0
        for c in self.vectorConstraints:
            c.computeConstraintsMatrixAndBounds(qpInput)
            m = c.getNConstraints()
            self.linearMatrix[n:n + m, :] = c.linearMatrix
            self.lowerBound[n:n + m] = c.lowerBound
            self.upperBound[n:n + m] = c.upperBound
            n += m

    def update(self, qpInput):  # :150-194
        if self.firstUpdate:
            self.hessian[:] = 0
        self.gradient[:] = 0
        for c in self.vectorCosts:
            c.computeHessianAndGradient(qpInput)
            if self.firstUpdate:
                self.hessian += c.hessian
            self.gradient += c.gradient
        self.firstUpdate = False
        self._gather_constraints(qpInput)
        return True

    def solve(self):  # :196-298
        if self.qp_solver is None:
            z, y, info = solve_qp_exact(self.hessian, self.gradient, self.linearMatrix,
                                        self.lowerBound, self.upperBound)
            self.solveInfo = info
            self.statusQPProblem = self.SOLVED
        else:
            z, status, info = self.qp_solver(self.hessian, self.gradient, self.linearMatrix,
                                             self.lowerBound, self.upperBound)
            self.solveInfo = info
            self.statusQPProblem = status
        self.outputQP = z
        return True

    def getSolution(self): return self.outputQP
    def getHessian(self): return self.hessian
    def getGradient(self): return self.gradient
    def getLinearConstraintMatrix(self): return self.linearMatrix
    def getLowerBound(self): return self.lowerBound
    def getUpperBound(self): return self.upperBound
    def getQPProblemStatus(self): return self.statusQPProblem
    def getNOptimizationVariables(self): return self.nVar
    def getNConstraints(self): return self.nConstraints


class VariableSamplingMPC(IMPCProblem):
    """MPC/src/variableSamplingMPC/variableSamplingMPC.cpp:7-227."""

    def setCostAndConstraints(self, p, qpInput):  # :7-86
        self.controlledJoints = list(p["controlledJoints"])
        self.nCtrlJoints = len(self.controlledJoints)
        assert self.nCtrlJoints == N_JOINTS
        self.nIter, self.nIterSmall, self.ctrlHorizon = p["nIter"], p["nIterSmall"], p["controlHorizon"]
        self.robot = qpInput.getRobot()
        self.nJets = self.robot.getNJets()
        self.jetModel = qpInput.getJetModel()
        self.nStates = rpyErrorIdx[2] + 1
        self.nInput = self.nCtrlJoints + self.nJets
        self.nVar = self.nStates * (self.nIter + 1) + self.nCtrlJoints * self.ctrlHorizon \
            + self.nJets * (self.ctrlHorizon - self.nIterSmall + 1)
        self.jointSelectorVector = [i for name in self.controlledJoints
                                    for i in range(self.robot.getNJoints())
                                    if name == self.robot.joint_names[i]]
        self.jointsPositionReference = self.robot.joint_pos.copy()
        self.deltaJointsPositionReference = np.zeros(self.nCtrlJoints)
        self.thrustReference = np.zeros(self.nJets)
        self.thrustDotReference = np.zeros(self.nJets)
        self.throttleReference = np.zeros(self.nJets)
        self.finalState = np.zeros(self.nStates)
        nV, nS, nJ, nT = self.nVar, self.nStates, self.nCtrlJoints, self.nJets
        self.vectorCosts = [ReferenceTrackingCost(nV, nS, self.nIter),
                            RegualarizationCost(nV, nS, nJ, nT),
                            ThrottleInitialValueCost(nV, nS, nJ, nT),
                            JointPositionRegularizationCost(nV, nS, nJ)]
        self.vectorConstraints = [ConstraintSystemDynamicVS(nV, nS, nJ, nT, self.nIter),
                                  ConstraintInitialState(nS, nV),
                                  ThrottleConstraint(nV, nS, self.nIter, self.nIterSmall)]
        if p.get("jointPos_max") is not None and p.get("jointPos_min") is not None:
            # optional extension (not in the shipped problem): per-block joint-limit rows, see JointPositionConstraint
            self.vectorConstraints.append(JointPositionConstraint(nV, nS, nJ, self.nIter))
        return True

    def solveMPC(self):  # :88-112
        self.solve()
        if self.getQPProblemStatus() == self.SOLVED:
            z = self.getSolution()
            nS, nJ, nT = self.nStates, self.nCtrlJoints, self.nJets
            states = z[:nS * (self.nIter + 1)]
            inputs = z[nS * (self.nIter + 1):]
            self.statesSolution, self.inputSolution = states, inputs
            self.deltaJointsPositionReference = inputs[0:nJ].copy()
            self.throttleReference = inputs[nJ * self.ctrlHorizon: nJ * self.ctrlHorizon + nT].copy()
            self.thrustReference = states[nS + thrustIdx[0]: nS + thrustIdx[0] + nT].copy()
            self.thrustDotReference = states[nS + thrustDotIdx[0]: nS + thrustDotIdx[0] + nT].copy()
            self.finalState = states[-nS:].copy()
            for i, j in enumerate(self.jointSelectorVector):
                self.jointsPositionReference[j] += self.deltaJointsPositionReference[i]
        return True

    def getJointsReferencePosition(self): return self.jointsPositionReference.copy()

    def getThrottleReference(self):  # :138-151
        return np.array([self.jetModel.destandardizeThrottle_u2T(v) for v in self.throttleReference])

    def getThrustReference(self): return self.thrustReference.copy()
    def getThrustDotReference(self): return self.thrustDotReference.copy()
    def getFinalCoMPosition(self): return self.finalState[0:3].copy()
    def getFinalLinMom(self): return self.finalState[3:6].copy()
    def getFinalRPY(self): return self.finalState[6:9].copy()
    def getFinalAngMom(self): return self.finalState[9:12].copy()
    def getNStatesMPC(self): return float(self.nStates)
    def getNInputMPC(self): return float(self.nInput)
