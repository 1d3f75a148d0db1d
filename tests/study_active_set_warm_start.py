#!/usr/bin/env python
"""CPU study (not a test; imports the oracle, hence under tests/): exchange pivots of the condensed kernels' active set on
the bench workload of every horizon, cold (inverse of the reduced Hessian + dual iterations: what the CUDA kernels do today)
against tools/condensed_model.box_qp_pivot_warm started from the working set of the previous state / from the all-lower vertex.
The reduced throttle QP is obtained from the oracle's dense QP by eliminating states and joint increments through the
equality rows.  Writes the table of profiles/r01i_active_set_warm_start.md to stdout.
usage: python tests/study_active_set_warm_start.py [instances per horizon]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import condensed_model as cm                                   # noqa: E402
from helpers import load_trajectories, pkg                     # noqa: E402
from oracle_driver import OracleInstance                       # noqa: E402


def build(params, i, seed):
    syn = pkg("synthetic")
    nom = syn.make_states(i + 1, perturbed=False)
    per = syn.make_states(i + 1, seed=seed, perturbed=True, near_bound_fraction=0.1)
    o = OracleInstance(nom, i, params=params, trajectories=load_trajectories())
    o.update(per)
    return o


def reduced_throttle_qp(o):
    """(H, g, lo, up) over the throttle variables that are not pinned: y = (x, dq) eliminated through the KKT system of the
    equality rows (dynamics + initial state), y = y0 + Y v."""
    m = o.mpc
    P, q, A, l, u = m.hessian, m.gradient, m.linearMatrix, m.lowerBound, m.upperBound
    N, Nc = o.params["nIter"], o.params["controlHorizon"]
    ny, neq = 26 * (N + 1) + 8 * Nc, 26 * N + 26
    nv = m.nVar - ny
    assert np.abs(u[:neq] - l[:neq]).max() == 0 and np.abs(P[:ny, ny:]).max() == 0
    Ey, Ev, Pyy = A[:neq, :ny], A[:neq, ny:], P[:ny, :ny]
    K = np.block([[Pyy, Ey.T], [Ey, np.zeros((neq, neq))]])
    sol = np.linalg.solve(K, np.column_stack([np.concatenate([-q[:ny], l[:neq]]), np.vstack([np.zeros((ny, nv)), -Ev])]))
    y0, Y = sol[:ny, 0], sol[:ny, 1:]
    H = P[ny:, ny:] + Y.T @ Pyy @ Y
    H = 0.5 * (H + H.T)
    g = q[ny:] + Y.T @ (Pyy @ y0 + q[:ny])
    lo, up = np.full(nv, -np.inf), np.full(nv, np.inf)
    for r in range(neq, A.shape[0]):
        nz = np.flatnonzero(A[r, ny:])
        if nz.size == 1:
            lo[nz[0]], up[nz[0]] = l[r] / A[r, ny + nz[0]], u[r] / A[r, ny + nz[0]]
    fixed, fr = np.flatnonzero(lo == up), np.flatnonzero(lo != up)
    return H[np.ix_(fr, fr)], g[fr] + H[np.ix_(fr, fixed)] @ lo[fixed], float(lo[fr][0]), float(up[fr][0])


def cold(H, g, lo, up):
    n = H.shape[0]
    T = np.array(H, float)
    for q in range(n):
        cm.exchange_pivot(T, q)
    vv, act, lam = -T @ g, np.zeros(n, int), np.zeros(n)
    status, it = cm._dual_pivot_loop(T, vv, act, lam, lo, up, 40 * n, 1e-10)
    assert status == 0
    return vv, act, n + it


if __name__ == "__main__":
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    cases = [("1x (17, 7, 12)", None), ("2x (34, 14, 24)", dict(nIter=34, nIterSmall=14, controlHorizon=24)),
             ("3x (51, 14, 36)", dict(nIter=51, nIterSmall=14, controlHorizon=36)),
             ("4x (68, 14, 48)", dict(nIter=68, nIterSmall=14, controlHorizon=48))]
    print("| horizon | instance | throttle variables | at a bound | pivots cold (inverse + dual) | warm, previous state's set | warm, all-lower vertex |")
    print("|---|---|---|---|---|---|---|")
    for name, params in cases:
        for i in range(n_inst):
            _, act_prev, _ = cold(*reduced_throttle_qp(build(params, i, 30)))      # "previous tick": another perturbed state
            H, g, lo, up = reduced_throttle_qp(build(params, i, 31))
            v, act, pc = cold(H, g, lo, up)
            w, _, s1, p1 = cm.box_qp_pivot_warm(H, g, lo, up, act_prev, max_iter=40 * H.shape[0])
            w2, _, s2, p2 = cm.box_qp_pivot_warm(H, g, lo, up, -np.ones(H.shape[0], int), max_iter=40 * H.shape[0])
            assert s1 == 0 and s2 == 0 and np.abs(w - v).max() < 1e-8 and np.abs(w2 - v).max() < 1e-8
            print(f"| {name} | {i} | {H.shape[0]} | {int((act != 0).sum())} | {pc} | {p1} | {p2} |", flush=True)
