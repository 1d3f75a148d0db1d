"""Per-instance, per-tick input pack: the values ``VariableSamplingMPC::update(QPInput&)`` reads.

The pack is a structure-of-arrays FP64 matrix of shape ``(PACK_DOUBLES, B)`` — row = one scalar of
one field, column = MPC instance — so that a thread-per-instance kernel reads it fully coalesced.
Row offsets here must match ``include/vsmpc.h`` (``VSMPC_PK_*``).

Each field is the output of a getter the reference calls on the path:

=================  ====  ==========================================================================
field              size  reference read
=================  ====  ==========================================================================
wRb                9     Robot::getBasePose().getRotation()         systemDynamicsVSMPC.cpp:107,324
omega_world        3     getBaseVel().getAngularVec3()              :108,325
rpy                3     getRotation().asRPY()                      :132; constraintsVSMPC.cpp:236-246
mass               1     getTotalMass() (float-rounded)             :297,308; Robot.h:338
gravity            3     getGravity()                               :309
M_b                36    getMassMatrix().block(0,0,6,6) row-major   :116,128-130; costsVSMPC.cpp:274
base_pos           3     getBasePose().getPosition()                :112
p_com              3     getPositionCoM()                           :111; constraintsVSMPC.cpp:210
momentum_body      6     getMomentum(true)                          constraintsVSMPC.cpp:211-214
A_mom_body         24    getMatrixAmomJets(true) 6x4 row-major      :93,304
jet_axes           12    getMatrixOfJetAxes()[i] (world) 4x3        :173,342
jet_arms           12    getMatrixOfJetArms()[i] (world) 4x3        :179
J_rel_ang          96    getRelativeJacobianJetsBodyFrame()[i].bottomRows(3), controlled-joint
                         columns, 4x3x8                             :183,343
J_jet_lin          96    getJacobian(frame).topRightCorner(3,nJ), controlled-joint columns :211
J_com              24    getJacobianCoM().topRightCorner(3,nJ), controlled-joint columns   :220
thrust             4     getJetThrusts()                            :170,403; constraintsVSMPC.cpp:217
thrust_dot_est     4     QPInput::getEstimatedThrustDot()           :404
thrust_des         4     QPInput::getThrustDesMPC()                 :415
thrust_dot_des     4     QPInput::getThrustDotDesMPC()              :415
throttle_prev      4     QPInput::getThrottleMPC()                  :411; costsVSMPC.cpp:484
q_cmd              8     QPInput::getOutputQPJointsPosition()[controlled]  costsVSMPC.cpp:581
=================  ====  ==========================================================================
"""
from __future__ import annotations

import numpy as np

PACK_FIELDS = [
    ("wRb", 9), ("omega_world", 3), ("rpy", 3), ("mass", 1), ("gravity", 3),
    ("M_b", 36), ("base_pos", 3), ("p_com", 3), ("momentum_body", 6), ("A_mom_body", 24),
    ("jet_axes", 12), ("jet_arms", 12), ("J_rel_ang", 96), ("J_jet_lin", 96), ("J_com", 24),
    ("thrust", 4), ("thrust_dot_est", 4), ("thrust_des", 4), ("thrust_dot_des", 4),
    ("throttle_prev", 4), ("q_cmd", 8),
]
PACK_OFFSETS = {}
_o = 0
for _n, _s in PACK_FIELDS:
    PACK_OFFSETS[_n] = (_o, _s)
    _o += _s
PACK_DOUBLES = _o  # 359

# controlled joints = columns 3..10 of the 23-joint axes list (src/config/robot.toml:3-27)
DEFAULT_JOINT_SELECTOR = list(range(3, 11))


def build_pack(state: dict, sel=None) -> np.ndarray:
    """``state``: getter-level batch dict (see synthetic.py; arrays with leading dim B, Jacobians with
    all nJ joint columns).  Returns the (PACK_DOUBLES, B) FP64 SoA matrix."""
    sel = DEFAULT_JOINT_SELECTOR if sel is None else list(sel)
    B = state["wRb"].shape[0]
    out = np.empty((PACK_DOUBLES, B), dtype=np.float64)

    def put(name, arr):
        o, s = PACK_OFFSETS[name]
        a = np.asarray(arr, dtype=np.float64).reshape(B, -1)
        if a.shape[1] != s:
            raise ValueError(f"pack field {name}: expected {s} scalars, got {a.shape[1]}")
        out[o:o + s, :] = a.T

    put("wRb", state["wRb"])
    put("omega_world", state["omega_world"])
    put("rpy", state["rpy"])
    put("mass", state["mass"])
    put("gravity", state["gravity"])
    put("M_b", state["M_b"])
    put("base_pos", state["base_pos"])
    put("p_com", state["p_com"])
    put("momentum_body", state["momentum_body"])
    put("A_mom_body", state["A_mom_body"])
    put("jet_axes", state["jet_axes"])
    put("jet_arms", state["jet_arms"])
    put("J_rel_ang", state["J_rel_body"][:, :, 3:6, :][..., sel])
    put("J_jet_lin", state["J_jet_lin"][..., sel])
    put("J_com", state["J_com"][..., sel])
    put("thrust", state["thrust"])
    put("thrust_dot_est", state["thrust_dot_est"])
    put("thrust_des", state["thrust_des"])
    put("thrust_dot_des", state["thrust_dot_des"])
    put("throttle_prev", state["throttle_prev"])
    put("q_cmd", state["q_cmd"][:, sel])
    return out
