#!/usr/bin/env python
"""GPU-box study (not a test; imports the oracle, hence under tests/): who is right about the closed loops the QP kernel
gives up on?  Runs the Monte Carlo sweep of tools/montecarlo_sweep.py (BASELINE configs[2] x configs[4]), then ONE more
tick without the CUDA graph, and for every instance whose status is != 0 on that tick hands the QP the library itself
assembled for it — IMPCProblem::getHessian / getGradient / getLinearConstraintMatrix / getLower/UpperBound — to the
oracle's exact sparse-KKT solver (oracle/vsmpc_oracle.solve_qp_exact, the arbiter of "matched KKT tolerance") and, when
oracle/_ref is present, compares the assembled gradient / bounds with nothing else (the compiled reference needs a Robot).
Writes the table of profiles/r02_nonsolved_adjudication.md to gpurun_out/.
usage: python tests/adjudicate_nonsolved.py [B] [ticks]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench                                                   # noqa: E402
from oracle import vsmpc_oracle as O                           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 199
MAX_CASES = 48
bat, ro, syn, cfg, L = bench.pkg("batched"), bench.pkg("rollout"), bench.pkg("synthetic"), bench.pkg("config"), bench.pkg("_lib")
rb = syn.SyntheticRobot()
g = np.random.default_rng(20251002)
ms, isc = g.uniform(0.9, 1.1, B), g.uniform(0.8, 1.2, B)
st = syn.make_states(B, seed=20251002, perturbed=True, near_bound_fraction=0.0, mass_scale=ms, inertia_scale=isc)
hover = (rb.mass * ms * 9.81 / 4.0)[:, None]
st["thrust"] = hover + g.normal(0, 8.0, (B, 4)); st["thrust_des"] = st["thrust"].copy()
st["thrust_dot_est"] = g.normal(0, 5.0, (B, 4)); st["thrust_dot_des"] = np.zeros((B, 4))
st["throttle_prev"] = np.full((B, 4), 76.0) + g.normal(0, 3.0, (B, 4))
st["momentum_body"] *= 0.2
st["q_cmd"] = np.tile(rb.joint_pos0, (B, 1))
dT = g.normal(0, 10.0, (B, 4))
coeff = np.tile(np.asarray(cfg.JET_COEFF), (B, 1)); norm = np.tile(np.asarray(cfg.JET_NORM), (B, 1))
coeff[:, 1] *= g.uniform(0.9, 1.1, B); coeff[:, 2] *= g.uniform(0.9, 1.1, B)
norm[:, 0] *= g.uniform(0.9, 1.1, B); norm[:, 1] *= g.uniform(0.9, 1.1, B)
tmin, tmax = g.uniform(0.0, 20.0, B), g.uniform(80.0, 100.0, B)
trj = dict(bench.load_traj()); trj["alphaGravity"] = np.ones_like(trj["alphaGravity"])
mpc = bat.BatchedVSMPC(B, None, trj, device=0, full_solution=True)
mpc.set_instance_params(coeff, norm, tmin, tmax)
loop = ro.BatchedRollout(mpc, rb)
loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT, phase0=(np.arange(B) % 20).astype(np.int32))
t0 = time.perf_counter()
loop.run(ticks)
pack = loop.pack()                     # the pack the next tick solves on
loop.run(1, use_graph=False)           # ... this tick: its QP data stays on the device for the getters
t_run = time.perf_counter() - t0
out, status = mpc.get_output()
z = mpc.getSolution()
q, l, u = mpc.get_qp_vectors()
piv = mpc.get_pivot_counts()
bad = np.flatnonzero(status != 0)
H = mpc.getHessian(0)
lines = [f"# r02 — adjudication of the closed loops the QP kernel does not solve ({B} loops x {ticks + 1} ticks, Monte Carlo sweep "
         f"of configs[2] x configs[4], one B200)\n",
         f"* status != 0 on tick {ticks + 1}: {bad.size} of {B} loops (status 1: {(status == 1).sum()}, status 2: {(status == 2).sum()}); "
         f"exchange pivots on that tick: mean {piv.mean():.1f}, max {piv.max()}; wall time {t_run:.2f} s",
         "* each case below: the QP exactly as the library assembled it for that tick (getHessian / getGradient / "
         "getLinearConstraintMatrix / bounds) solved by the oracle's exact sparse-KKT active-set solver\n",
         "| loop | status | finite QP data | max abs x0 | pitch [rad] | 1/cos(pitch) | max abs A entry | oracle verdict | oracle KKT (stationarity / primal) | "
         "max abs z_oracle |", "|---|---|---|---|---|---|---|---|---|---|"]
n_exact_ok = n_exact_fail = n_nonfinite = 0
for i in bad[:MAX_CASES]:
    A = mpc.getLinearConstraintMatrix(int(i))
    fin = bool(np.isfinite(A).all() and np.isfinite(q[i]).all() and np.isfinite(l[i]).all() and np.isfinite(u[i]).all())
    x0 = l[i, 442:468]
    pitch = pack[13, i]                                # VSMPC_PK_RPY + 1
    verdict, kkt, zmax = "-", "-", "-"
    if not fin:
        n_nonfinite += 1
        verdict = "QP data not finite: no solver can solve it"
    else:
        try:
            zo, yo, info = O.solve_qp_exact(H, q[i], A, l[i], u[i])
            ok = bool(np.isfinite(zo).all()) and info["stationarity"] < 1e-6 * max(1.0, np.abs(q[i]).max()) and info["primal"] < 1e-6
            verdict = "exact solver succeeds" if ok else "exact solver returns a point that fails its own KKT check"
            kkt = f"{info['stationarity']:.1e} / {info['primal']:.1e}"
            zmax = f"{np.abs(zo).max():.3e}"
            n_exact_ok += ok
            n_exact_fail += (not ok)
        except Exception as e:       # singular KKT system, cycling, ...
            verdict = f"exact solver fails: {type(e).__name__}: {str(e)[:60]}"
            n_exact_fail += 1
    with np.errstate(all="ignore"):
        lines.append(f"| {i} | {status[i]} | {fin} | {np.abs(x0).max():.3e} | {pitch:.3f} | {1 / np.cos(pitch):.2e} | "
                     f"{np.abs(A[np.isfinite(A)]).max():.2e} | {verdict} | {kkt} | {zmax} |")
lines += ["", f"Summary of the {min(bad.size, MAX_CASES)} adjudicated cases: QP data not finite {n_nonfinite}; exact solver succeeds "
          f"{n_exact_ok}; exact solver fails / fails its KKT check {n_exact_fail}."]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r02_nonsolved_adjudication.md"), "w").write("\n".join(lines) + "\n")
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "r02_nonsolved_packs.npz"), idx=bad[:MAX_CASES], pack=pack[:, bad[:MAX_CASES]],
                    q=q[bad[:MAX_CASES]], l=l[bad[:MAX_CASES]], u=u[bad[:MAX_CASES]], status=status[bad[:MAX_CASES]])
print("\n".join(lines))
