#!/usr/bin/env python
"""Closed-loop experiments on the GPU box (SURVEY §8d Config 1 and §8c), written to gpurun_out/closed_loop_report.md:
  1. one closed loop (B = 1), reference configuration and trajectory fixtures, surrogate plant, 4000 ticks (20 s):
     per-tick latency of the device-resident loop (vsmpc_rollout_run(1 tick) from the host) p50 / p99, tracking error;
  2. long-horizon parity: device loop vs the oracle MPC in closed loop over the oracle's surrogate plant, N ticks,
     maximum position / attitude / thrust / throttle deviation (SURVEY asks <= 1e-4 m, 1e-4 rad over 10 s)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from helpers import load_trajectories
from oracle_driver import oracle_trajectories_to_product
from test_rollout import make_case, oracle_plant
from oracle.plant_surrogate import SurrogateLoop

bat, ro, syn = bench.pkg("batched"), bench.pkg("rollout"), bench.pkg("synthetic")
n_parity = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
out = ["# closed-loop experiments (surrogate plant, DESIGN.md §10)\n"]

# ---- 1. single loop, 4000 ticks --------------------------------------------------------------------------------
rb = syn.SyntheticRobot()
st = syn.make_states(1, perturbed=False)
hover = rb.mass * 9.81 / 4.0
st["thrust"][:] = hover; st["thrust_des"][:] = hover; st["throttle_prev"][:] = 76.0     # in flight, near the hover throttle
# hover scenario: the take-off ramp of alphaGravity (0.08 -> 1 over 20 s, during which the ground carries the robot and
# which the surrogate has no contact model for) is replaced by alphaGravity = 1; the CoM / RPY reference is the fixture's
trj = dict(bench.load_traj())
trj["alphaGravity"] = np.ones_like(trj["alphaGravity"])
mpc = bat.BatchedVSMPC(1, None, trj)
loop = ro.BatchedRollout(mpc, rb)
loop.init(st)
loop.run(20)
ts, rec = [], []
for t in range(4000):
    t0 = time.perf_counter()
    r = loop.run(1, record_every=1)
    ts.append(time.perf_counter() - t0)
    rec.append(r[0, 0])
rec = np.array(rec); ts = np.array(ts) * 1e3
out.append("## 1. one closed loop, reference configuration, 4000 ticks (20 s of controller time)\n")
out.append(f"* per-tick latency of the device-resident loop (plant + linearise + QP kernels, one `vsmpc_rollout_run(1)` call per tick, "
           f"host wall clock incl. the 128-byte record read-back): p50 {np.percentile(ts,50):.3f} ms, p99 {np.percentile(ts,99):.3f} ms, "
           f"max {ts.max():.3f} ms (controller period 5 ms; reference poster: 2.18 ms mean)")
out.append(f"* solver status: {int((rec[:,14]==0).sum())} / 4000 ticks solved")
err = np.abs(rec[:, 0:3] - st["p_com"][0])
out.append(f"* hover (alphaGravity = 1, stationary CoM / RPY reference of the fixture's first 20 s): max CoM deviation {err.max()*1e3:.1f} mm "
           f"(at t = {0.005*(1+int(err.max(axis=1).argmax())):.2f} s), over the last 5 s {err[-1000:].max()*1e3:.2f} mm; |rpy| max "
           f"{np.abs(rec[:,3:6]).max()*1e3:.1f} mrad, over the last 5 s {np.abs(rec[-1000:,3:6]).max()*1e3:.2f} mrad; thrust "
           f"{rec[:,6:10].min():.1f} .. {rec[:,6:10].max():.1f} N per jet (hover {hover:.1f} N); throttle {rec[:,10:14].min():.1f} .. {rec[:,10:14].max():.1f} %\n")
mpc.close()

# ---- 2. long parity run ---------------------------------------------------------------------------------------------
B = 2
rbp, stp, ms, isc, dT = make_case(B, seed=11)
dT = 0.3 * dT                                   # constant thrust disturbances ~ N(0, 3 N) per jet
stp["momentum_body"] = 0.2 * stp["momentum_body"]
for i in range(B):
    stp["omega_world"][i] = stp["wRb"][i] @ np.linalg.solve(rbp.I_body * isc[i], stp["momentum_body"][i, 3:])
traj = load_trajectories()
traj["TRAJECTORY_MANAGER"]["arrays"]["alphaGravity"] = np.ones_like(traj["TRAJECTORY_MANAGER"]["arrays"]["alphaGravity"])
mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
loop = ro.BatchedRollout(mpc, rbp)
loop.init(stp, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
t0 = time.perf_counter()
rec = loop.run(n_parity, record_every=1)
t_gpu = time.perf_counter() - t0
worst = np.zeros(4)
t0 = time.perf_counter()
for i in range(B):
    o = SurrogateLoop(oracle_plant(rbp, stp, i, ms[i], isc[i], dT[i]), trajectories=traj)
    for t in range(n_parity):
        r = o.tick()
        worst[0] = max(worst[0], np.abs(rec[t, i, 0:3] - r[0:3]).max())
        worst[1] = max(worst[1], np.abs(rec[t, i, 3:6] - r[3:6]).max())
        worst[2] = max(worst[2], np.abs(rec[t, i, 6:10] - r[6:10]).max())
        worst[3] = max(worst[3], np.abs(rec[t, i, 10:14] - r[10:14]).max())
t_cpu = time.perf_counter() - t0
out.append(f"## 2. device loop vs oracle loop, {B} instances x {n_parity} ticks ({n_parity*0.005:.1f} s), in flight (alphaGravity = 1), "
           f"mass / inertia scaled, constant thrust disturbances, perturbed initial state\n")
out.append(f"* excursion of the loops themselves: max |CoM - start| {np.abs(rec[:,:,0:3]-stp['p_com'][None]).max():.3f} m, max |rpy| {np.abs(rec[:,:,3:6]).max():.3f} rad, "
           f"final |CoM - start| {np.abs(rec[-1,:,0:3]-stp['p_com']).max():.3f} m")
out.append(f"* max |CoM position difference| {worst[0]:.3e} m, max |RPY difference| {worst[1]:.3e} rad "
           f"(bound asked by SURVEY §8c: 1e-4 m / 1e-4 rad over 10 s)")
out.append(f"* max |thrust difference| {worst[2]:.3e} N, max |throttle difference| {worst[3]:.3e} %")
out.append(f"* all ticks solved on the device: {bool((rec[:,:,14]==0).all())}; wall time device loop {t_gpu:.2f} s (incl. record copy), oracle loop {t_cpu:.1f} s")
open(os.path.join(ROOT, "gpurun_out", "closed_loop_report.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
