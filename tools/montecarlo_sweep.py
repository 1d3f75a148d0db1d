#!/usr/bin/env python
"""BASELINE.json configs[2] + configs[4] at full size on one GPU: a Monte Carlo closed-loop sweep of B instances
(default 65 536) over initial states, constant thrust disturbances and per-instance model parameters (jet
coefficients / normalisation, mass, inertia, throttle limits), 200 controller ticks each, entirely on the device.
Writes gpurun_out/montecarlo_sweep.md.  Under torchrun every rank runs its own shard (seed = base + rank)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bat, ro, syn, cfg = bench.pkg("batched"), bench.pkg("rollout"), bench.pkg("synthetic"), bench.pkg("config")
rb = syn.SyntheticRobot()
g = np.random.default_rng(20251002 + rank)
t0 = time.perf_counter()
ms, isc = g.uniform(0.9, 1.1, B), g.uniform(0.8, 1.2, B)
st = syn.make_states(B, seed=20251002 + rank, perturbed=True, near_bound_fraction=0.0, mass_scale=ms, inertia_scale=isc)
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
t_gen = time.perf_counter() - t0
trj = dict(bench.load_traj()); trj["alphaGravity"] = np.ones_like(trj["alphaGravity"])      # in flight
mpc = bat.BatchedVSMPC(B, None, trj, device=local)
mpc.set_instance_params(coeff, norm, tmin, tmax)
loop = ro.BatchedRollout(mpc, rb)
loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT, phase0=(np.arange(B) % 20).astype(np.int32))
loop.run(2)
torch.cuda.synchronize()
t0 = time.perf_counter()
rec = loop.run(ticks, record_every=ticks // 4)
t_run = time.perf_counter() - t0
_, status = mpc.get_output()
err = np.linalg.norm(rec[:, :, 0:3] - st["p_com"][None], axis=2)        # (4, B)
att = np.abs(rec[:, :, 3:6]).max(axis=2)
lines = [f"# Monte Carlo closed-loop sweep (configs[2] x configs[4]), rank {rank}: {B} instances x {ticks} ticks on one B200\n",
         f"* per-instance: initial state (SURVEY §8d Config 2 perturbations, in flight), constant thrust disturbance N(0, 10 N) per jet, "
         f"mass x U(0.9, 1.1), inertia x U(0.8, 1.2), jet c1, c2, mu_T, sigma_T x U(0.9, 1.1), throttleMin U(0, 20), throttleMax U(80, 100); "
         f"20-tick throttle phases staggered; surrogate plant (DESIGN.md §10)",
         f"* device time for {ticks} ticks (+ 4 record rows per instance copied back): {t_run:.3f} s = {B*ticks/t_run/1e6:.2f} M closed-loop solves/s "
         f"({t_run/ticks*1e3:.2f} ms per tick of {B} instances); host-side generation of the sweep: {t_gen:.1f} s",
         f"* solved at the last tick: {int((status==0).sum())} / {B}; ticks with status != 0 among the recorded rows: {int((rec[:,:,14]!=0).sum())}",
         "", "| t [s] | CoM distance from start: median / 95 % / max [m] | max |rpy|: median / 95 % / max [rad] |", "|---|---|---|"]
for k in range(rec.shape[0]):
    fin = np.isfinite(err[k]) & np.isfinite(att[k])
    e_, a_ = err[k][fin], att[k][fin]
    lines.append(f"| {(k+1)*(ticks//4)*0.005:.2f} | {np.median(e_):.3f} / {np.percentile(e_,95):.3f} / {e_.max():.3f} | "
                 f"{np.median(a_):.3f} / {np.percentile(a_,95):.3f} / {a_.max():.3f} |")
    if not fin.all():
        bad = ~fin
        lines.append(f"| | {int(bad.sum())} loops with a non-finite plant state (left out of the row above); "
                     f"status != 0 at that row for {int((rec[k, bad, 14] != 0).sum())} of them, at the last tick for "
                     f"{int((status[bad] != 0).sum())} | |")
open(os.path.join(ROOT, "gpurun_out", f"montecarlo_sweep_rank{rank}.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
