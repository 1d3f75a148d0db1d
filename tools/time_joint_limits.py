#!/usr/bin/env python
"""Development: cost of the optional joint-limit rows at the reference horizon (BASELINE configs[4] share of one GPU).
Three handles on the same states: no joint-limit rows; rows on with boxes nobody reaches; rows on with the tight boxes of
tests/test_gpu_joint_limits.py (an active joint bound in a good part of the solves)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
bat, syn, pack = bench.pkg("batched"), bench.pkg("synthetic"), bench.pkg("pack")
JMIN = [-30.0, 5.0, 20.0, 5.0, -30.0, 5.0, 20.0, 5.0]
JMAX = [-8.0, 30.0, 42.0, 28.0, -8.0, 30.0, 42.0, 28.0]
WIDE = dict(jointPos_min=[-400.0] * 8, jointPos_max=[400.0] * 8)
for B in (int(a) for a in (sys.argv[1:] or ["2048"])):
    nom = syn.make_states(B, perturbed=False)
    pers = [syn.make_states(B, seed=300 + j, perturbed=True, near_bound_fraction=0.3) for j in range(2)]
    nom_pack, packs = pack.build_pack(nom), [pack.build_pack(p) for p in pers]
    jp = np.ascontiguousarray(nom["joint_pos"][:, pack.DEFAULT_JOINT_SELECTOR].T)
    TIGHT = dict(jointPos_min=JMIN, jointPos_max=JMAX)
    for name, params, warm, nsets in (("no joint-limit rows", None, 1, 2), ("rows on, never active", WIDE, 1, 2),
                                      ("tight boxes, cold working set", TIGHT, 0, 2),
                                      ("tight boxes, warm, new states", TIGHT, 1, 2),
                                      ("tight boxes, warm, same state", TIGHT, 1, 1)):
        mpc = bat.BatchedVSMPC(B, params, bench.load_traj())
        mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
        mpc.set_warm_start(warm)
        for j in range(3):
            mpc.update_pack(packs[j % nsets]); mpc.solveMPC()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 10
        for j in range(n):
            mpc.update_pack(packs[j % nsets]); mpc.solveMPC()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        nf, ns = mpc.get_counts()
        _, status = mpc.get_output()
        hist = np.bincount(nf, minlength=9)[:9]
        print(f"B={B} {name:30s}: {dt*1e3:8.3f} ms per tick ({B/dt/1e6:.3f} M solves/s), solved {np.mean(status == 0):.4f}, "
              f"factorisations per solve: mean {nf.mean():.2f} max {nf.max()} histogram 1..8 {hist[1:].tolist()}")
        mpc.close()
