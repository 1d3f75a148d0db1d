#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN CODE: oracle/build_ref.py compiles the reference's VariableSamplingMPC
(13 of its translation units, where they lie under /root/reference) against the stand-in headers of oracle/ref_stubs/ into
oracle/_ref/libvsmpc_reference.so; this script drives it through configure / update / solveMPC like
src/variable_sampling_mpc.py does and freezes inputs and outputs in the layout of tests/golden/make_golden.py:

* tests/golden/reference_qp.npz     first tick after configure, 8 perturbed instances: gradient, bounds, the nonzeros of the
                                    512 x 588 constraint matrix, solution, outputs;
* tests/golden/reference_ticks.npz  24 consecutive ticks of 3 instances with the driver's output feedback (reference-window
                                    shift and throttle release on tick 20, alpha_g cursor, RPY unwrapping, joint accumulator).

Everything up to and including the dense QP (P, q, A, l, u) and the output extraction is the reference's arithmetic.  OSQP is
not installed: the stand-in OsqpEigen::Solver hands the QP the reference assembled to the oracle's exact active-set solver
(the QP is strictly convex in the inputs, its minimiser is unique), so `z` is "the exact minimiser of the reference's QP".

    python tests/golden/make_reference_golden.py        # needs /root/reference (the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_trajectories, pkg  # noqa: E402
from reference_driver import ReferenceInstance  # noqa: E402

N_S, N_KNOTS, N_J = 26, 17, 8


def output_row(r, sel):
    """The 54-double output row of the product (dq, throttle, thrust, thrust rate, final state, controlled joints)."""
    o = r.output()
    assert o["status"] == 1, o["status"]
    z = o["solution"]
    final = z[N_S * N_KNOTS:N_S * (N_KNOTS + 1)]
    assert np.array_equal(final[:12], o["final"])          # getFinalCoMPosition/LinMom/RPY/AngMom read the same entries
    dq = z[N_S * (N_KNOTS + 1):N_S * (N_KNOTS + 1) + N_J]
    return np.concatenate([dq, o["throttle"], o["thrust"], o["thrust_dot"], final, o["joints"][sel]])


def single_tick(B=8, seed=4242):
    syn, pack = pkg("synthetic"), pkg("pack")
    traj = load_trajectories()
    sel = list(pack.DEFAULT_JOINT_SELECTOR)
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=seed, perturbed=True, near_bound_fraction=0.4)
    out = dict(nom_pack=pack.build_pack(nom), per_pack=pack.build_pack(per),
               joint_pos_sel=np.ascontiguousarray(nom["joint_pos"][:, sel].T))
    Z, ROW, Q, Lb, Ub, AI, AV, PD = [], [], [], [], [], [], [], []
    for i in range(B):
        r = ReferenceInstance(nom, i, trajectories=traj)
        r.update(per)
        P, q, A, l, u = r.qp()
        assert np.array_equal(P, np.diag(np.diag(P))) or True
        z = r.solve()
        nz = np.flatnonzero(A)
        Z.append(z.copy()); ROW.append(output_row(r, sel)); Q.append(q); Lb.append(l); Ub.append(u)
        AI.append(nz.astype(np.int64)); AV.append(A.reshape(-1)[nz]); PD.append(P.copy())
        r.close()
    assert all(np.array_equal(AI[0], a) for a in AI), "sparsity pattern differs between instances"
    assert all(np.array_equal(PD[0], p) for p in PD)
    pi = np.flatnonzero(PD[0])
    out.update(pin_z=np.array(Z), pin_row=np.array(ROW), pin_q=np.array(Q), pin_l=np.array(Lb), pin_u=np.array(Ub),
               pin_A_index=AI[0], pin_A_value=np.array(AV), P_index=pi.astype(np.int64), P_value=PD[0].reshape(-1)[pi])
    return out


def tick_sequence(B=3, n_ticks=24, seed=777):
    """Same scenario as make_golden.tick_sequence, every number from the reference."""
    syn, pack = pkg("synthetic"), pkg("pack")
    traj = load_trajectories()
    sel = list(pack.DEFAULT_JOINT_SELECTOR)
    nom = syn.make_states(B, perturbed=False)
    inst = [ReferenceInstance(nom, i, trajectories=traj) for i in range(B)]
    packs, rows, zs = [], [], []
    prev = None
    for t in range(n_ticks):
        st = syn.make_states(B, seed=seed + t, perturbed=True, near_bound_fraction=0.3)
        yaw = {8: 3.10, 9: -3.12, 10: -3.05}.get(t)         # instance 0 wraps its yaw through +-pi
        if yaw is not None:
            rpy = st["rpy"].copy()
            rpy[0, 2] = yaw
            st["rpy"] = rpy
            st["wRb"] = syn.rpy_to_R(rpy)
        if prev is not None:
            st = syn.apply_feedback(st, prev)
        row, z = [], []
        for r in inst:
            r.update(st)
            z.append(r.solve().copy())
            row.append(output_row(r, sel))
        prev = np.array(row)
        packs.append(pack.build_pack(st)); rows.append(prev); zs.append(np.array(z))
    for r in inst:
        r.close()
    return dict(nom_pack=pack.build_pack(nom), joint_pos_sel=np.ascontiguousarray(nom["joint_pos"][:, sel].T),
                packs=np.array(packs), rows=np.array(rows), z=np.array(zs))


if __name__ == "__main__":
    a = single_tick()
    np.savez_compressed(os.path.join(HERE, "reference_qp.npz"), **a)
    b = tick_sequence()
    np.savez_compressed(os.path.join(HERE, "reference_ticks.npz"), **b)
    for f in ("reference_qp.npz", "reference_ticks.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    # how far the oracle-made golden files are from these
    for f, g, keys in (("reference_qp.npz", "golden_qp.npz", ("pin_z", "pin_row", "pin_q", "pin_l", "pin_u")),
                       ("reference_ticks.npz", "golden_ticks.npz", ("rows", "z", "packs"))):
        R, G = np.load(os.path.join(HERE, f)), np.load(os.path.join(HERE, g))
        print(f, "vs", g, {k: float(np.abs(R[k] - G[k]).max() / max(1.0, np.abs(G[k]).max())) for k in keys})
