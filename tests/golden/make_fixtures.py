#!/usr/bin/env python
"""Convert the reference's MAT-v7.3 trajectory fixtures to ``tests/golden/trajectories.npz``.

Runs ONLY in the build container (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures.py

Sources: /root/reference/src/trajectories/alphaGravity.mat (alphaGravity 1x351, fps 10) and
minimumJerkTrajectory.mat (positionCoM, velocityCoM, RPY, RPYDot 3x1481, fps 10), referenced from
/root/reference/src/config/vs_mcp_config.xml:34-40 and loaded by the reference through
TrajectoryManager::loadTrajectoryFromFile (utils/src/TrajectoryManager.cpp:67-140).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402
loadmat73 = importlib.import_module('paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200.mat73').loadmat73

REF = "/root/reference/src/trajectories"


def main():
    a = loadmat73(os.path.join(REF, "alphaGravity.mat"))
    m = loadmat73(os.path.join(REF, "minimumJerkTrajectory.mat"))
    out = dict(
        alpha_fps=np.array(int(a["fps"][0, 0])),
        alphaGravity=a["alphaGravity"],
        traj_fps=np.array(int(m["fps"][0, 0])),
        positionCoM=m["positionCoM"], velocityCoM=m["velocityCoM"],
        RPY=m["RPY"], RPYDot=m["RPYDot"],
    )
    assert out["alphaGravity"].shape == (1, 351)
    for k in ("positionCoM", "velocityCoM", "RPY", "RPYDot"):
        assert out[k].shape == (3, 1481), (k, out[k].shape)
    dst = os.path.join(ROOT, "tests", "golden", "trajectories.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
