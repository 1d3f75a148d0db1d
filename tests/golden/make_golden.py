#!/usr/bin/env python
"""Freeze golden vectors produced by the CPU oracle (oracle/vsmpc_oracle.py).

    python tests/golden/make_golden.py        # writes tests/golden/golden_qp.npz and golden_ticks.npz

The reference ships no golden vectors for this path ("parity unpinned"), so these files
pin the ORACLE: tests/test_oracle.py checks that it still reproduces them, tests/test_gpu_golden.py that
the CUDA path matches them.  Inputs (the packs) are stored next to the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_trajectories, pkg  # noqa: E402
from oracle_driver import OracleInstance  # noqa: E402


def single_tick(B=8, seed=4242):
    syn, pack = pkg("synthetic"), pkg("pack")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=seed, perturbed=True, near_bound_fraction=0.4)
    out = dict(nom_pack=pack.build_pack(nom), per_pack=pack.build_pack(per),
               joint_pos_sel=np.ascontiguousarray(nom["joint_pos"][:, pack.DEFAULT_JOINT_SELECTOR].T))
    for tag, free in (("pin", False), ("free", True)):
        Z, ROW, A, BJ, BT, C, Q, Lb, Ub, NA = [], [], [], [], [], [], [], [], [], []
        for i in range(B):
            o = OracleInstance(nom, i, trajectories=traj)
            if free:
                o.mpc.vectorConstraints[2].counter = 19
            o.update(per)
            z = o.solve()
            a, bj, bt, c, _ = o.dynamics()
            Z.append(z); ROW.append(o.output_row()); A.append(a.copy()); BJ.append(bj.copy()); BT.append(bt.copy())
            C.append(c.copy()); Q.append(o.mpc.gradient.copy()); Lb.append(o.mpc.lowerBound.copy())
            Ub.append(o.mpc.upperBound.copy()); NA.append(o.mpc.solveInfo["n_active"])
        for k, v in (("z", Z), ("row", ROW), ("A", A), ("BJ", BJ), ("BT", BT), ("c", C), ("q", Q), ("l", Lb),
                     ("u", Ub), ("n_active", NA)):
            out[f"{tag}_{k}"] = np.array(v)
    return out


def tick_sequence(B=3, n_ticks=24, seed=777):
    """24 consecutive ticks (covers the 20-tick reference shift / throttle release, the alpha_g cursor,
    RPY unwrapping across +-pi and the joint accumulator) with the driver's feedback of the outputs."""
    syn, pack = pkg("synthetic"), pkg("pack")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    inst = [OracleInstance(nom, i, trajectories=traj) for i in range(B)]
    packs, rows, zs = [], [], []
    prev = None
    for t in range(n_ticks):
        st = syn.make_states(B, seed=seed + t, perturbed=True, near_bound_fraction=0.3)
        # make instance 0 wrap its yaw through +-pi around ticks 8..10
        yaw = {8: 3.10, 9: -3.12, 10: -3.05}.get(t)
        if yaw is not None:
            rpy = st["rpy"].copy()
            rpy[0, 2] = yaw
            st2 = syn.make_states(B, seed=seed + t, perturbed=True, near_bound_fraction=0.3)
            st["rpy"] = rpy
            st["wRb"] = syn.rpy_to_R(rpy)
        if prev is not None:
            st = syn.apply_feedback(st, prev)
        row = []
        z = []
        for i, o in enumerate(inst):
            o.update(st)
            z.append(o.solve().copy())
            row.append(o.output_row())
        prev = np.array(row)
        packs.append(pack.build_pack(st)); rows.append(prev); zs.append(np.array(z))
    return dict(nom_pack=pack.build_pack(nom),
                joint_pos_sel=np.ascontiguousarray(nom["joint_pos"][:, pack.DEFAULT_JOINT_SELECTOR].T),
                packs=np.array(packs), rows=np.array(rows), z=np.array(zs))


if __name__ == "__main__":
    g = os.path.join(ROOT, "tests", "golden")
    a = single_tick()
    np.savez_compressed(os.path.join(g, "golden_qp.npz"), **a)
    print("golden_qp", {k: v.shape for k, v in a.items() if k.endswith("_z") or k.endswith("n_active")}, a["pin_n_active"], a["free_n_active"])
    b = tick_sequence()
    np.savez_compressed(os.path.join(g, "golden_ticks.npz"), **b)
    print("golden_ticks", b["packs"].shape, b["rows"].shape)
    for f in ("golden_qp.npz", "golden_ticks.npz"):
        print(f, os.path.getsize(os.path.join(g, f)), "bytes")
