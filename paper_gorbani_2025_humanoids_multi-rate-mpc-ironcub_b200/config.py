"""MPC parameters (group VS_MPC_CONFIG of the reference's src/config/vs_mcp_config.xml:5-44) and
trajectory fixtures as plain arrays (stand-in for TrajectoryManager's matio loader,
UT/src/TrajectoryManager.cpp:67-140)."""
from __future__ import annotations

import ast
import os
import xml.etree.ElementTree as ET

import numpy as np

JET_COEFF = [-4.64730485e-01, -8.13171858e+00, -6.19539230e+00, 6.61113140e-01, 1.67673231e+00,
             -4.83287064e-01, 8.77996617e+00, -1.01096376e+00, -5.86442286e-01, 5.19093322e-01,
             -4.23782666e-01, -1.45705257e+00, -7.83052261e-03]   # UT/src/JetModel.cpp:13-25
JET_NORM = [108.309, 65.793, 47.333, 31.483]                      # UT/src/JetModel.cpp:26


def default_params() -> dict:
    """Values of src/config/vs_mcp_config.xml (reference defaults)."""
    return dict(
        useJetDynamic=True, useEstimatedThrust=True,
        periodMPC=0.005, periodMPCLargeSteps=0.1, periodMPCSmallSteps=0.005,
        nIter=17, nIterSmall=7, controlHorizon=12,
        controlledJoints=["l_shoulder_pitch", "l_shoulder_roll", "l_shoulder_yaw", "l_elbow",
                          "r_shoulder_pitch", "r_shoulder_roll", "r_shoulder_yaw", "r_elbow"],
        jointsLambdaOption="unfiltered",
        weightCoMPos=[500.0, 500.0, 5000.0], weightCoMPosError=[25000.0, 25000.0, 50000.0],
        weightLinMom=[1.0, 1.0, 1.5], weightRPY=[1000.0, 1000.0, 1000.0],
        weightRPYError=[10000.0, 10000.0, 10000.0], weightAngMom=[80.0, 80.0, 80.0],
        weightDeltaJoint=[65000.0] * 8, weightThrottle=80000.0, weightInitialThrottle=80000.0,
        weightRegularizationJointPos=20.0, throttleMin=0.0, throttleMax=100.0,
    )


def _parse_value(text: str):
    t = text.strip()
    if t in ("true", "false"):
        return t == "true"
    if t.startswith("("):  # YARP list: (a b c) or ("a", "b")
        inner = t[1:-1].strip()
        if '"' in inner:
            return [s.strip().strip('"') for s in inner.split(",")]
        return [float(x) for x in inner.replace(",", " ").split()]
    if t.startswith('"'):
        return t.strip('"')
    try:
        return int(t)
    except ValueError:
        try:
            return float(t)
        except ValueError:
            return t


def read_xml_config(path: str, group: str = "VS_MPC_CONFIG") -> dict:
    """Reader for the reference's YARP-robotinterface XML (what flightCtrl.readXMLFile +
    get_group('VS_MPC_CONFIG') give the driver, src/variable_sampling_mpc.py:31-38).  Sub-groups
    become nested dicts."""
    root = ET.parse(path).getroot()

    def walk(node):
        d = {}
        for ch in node:
            if ch.tag == "param":
                d[ch.attrib["name"]] = _parse_value(ch.text or "")
            elif ch.tag == "group":
                d[ch.attrib["name"]] = walk(ch)
        return d

    for g in root.iter("group"):
        if g.attrib.get("name") == group:
            return walk(g)
    raise KeyError(group)


def load_trajectories_npz(path: str) -> dict:
    """Fixture arrays: alphaGravity (1,n) at alpha_fps; positionCoM/velocityCoM/RPY/RPYDot (3,m) at traj_fps."""
    d = np.load(path)
    return dict(alpha_fps=int(d["alpha_fps"]), alphaGravity=np.asarray(d["alphaGravity"], float),
                traj_fps=int(d["traj_fps"]), positionCoM=np.asarray(d["positionCoM"], float),
                velocityCoM=np.asarray(d["velocityCoM"], float), RPY=np.asarray(d["RPY"], float),
                RPYDot=np.asarray(d["RPYDot"], float))


def hover_trajectories(n: int = 8) -> dict:
    """Synthetic 'stay where you are' fixtures (alphaGravity = 1, zero reference motion)."""
    z = np.zeros((3, n))
    return dict(alpha_fps=10, alphaGravity=np.ones((1, n)), traj_fps=10, positionCoM=z, velocityCoM=z.copy(),
                RPY=z.copy(), RPYDot=z.copy())


def load_trajectories_mat(alpha_path: str, position_path: str) -> dict:
    """Load the two MAT-v7.3 files named by TRAJECTORY_MANAGER / POSITION_TRAJECTORY in the XML."""
    from .mat73 import loadmat73
    a, m = loadmat73(alpha_path), loadmat73(position_path)
    return dict(alpha_fps=int(a["fps"][0, 0]), alphaGravity=a["alphaGravity"], traj_fps=int(m["fps"][0, 0]),
                positionCoM=m["positionCoM"], velocityCoM=m["velocityCoM"], RPY=m["RPY"], RPYDot=m["RPYDot"])
